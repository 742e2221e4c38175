"""Generates the committed golden fixtures from the UNMODIFIED reference (run in the build
container only: needs /root/reference and oracle/_ref built by `make -C oracle ref`).

For every case the reference's own `db` / `qry` commands format the inputs, then
oracle/_ref/ref_probe (the reference Aligner driven stage by stage) dumps candidates, scores and
the final hit lists, and `ghostm aln -y 0/1/2` writes the output text.  Stored per case:
  db_<i>.seq.gz/.pos/.nam (the k-mer index is NOT stored: 4 MiB per chunk; tests rebuild it with
                         ghostm_b200.formats.build_index, which test_formats pins byte-for-byte
                         against the reference `db` output)
  db.meta.json           seed, chunk count, sum_residues, max_chunk_len
  q.inf q_<i>.inf/.seq/.nam
  dump.bin.gz            ref_probe stage dump
  out_y0.txt out_y1.txt out_y2.txt
  meta.json              aln options

usage: python tests/golden/make_golden.py [case ...]
"""
import gzip, json, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ghostm_b200 import formats, synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
TESTSET = "/root/reference/testset"


def run(*a, env=None):
    e = dict(os.environ)
    e.update(env or {})
    subprocess.check_call(list(a), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=e)


ONLY = set(sys.argv[1:])     # optional: regenerate just the named cases


def make_case(name, db_fasta, q_fasta, qry_args, db_args, aln_args, max_list=None, l1024=False):
    if ONLY and name not in ONLY:
        return
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    ghostm = os.path.join(REF, "ghostm_l1024" if l1024 else "ghostm")
    probe = os.path.join(REF, "ref_probe_l1024" if l1024 else "ref_probe")
    with tempfile.TemporaryDirectory() as tmp:
        run(ghostm, "db", "-i", db_fasta, "-o", f"{tmp}/db", *db_args)
        run(ghostm, "qry", "-i", q_fasta, "-o", f"{tmp}/q", *qry_args)
        env = {"GMPROBE_MAX_LIST_LENGTH": str(max_list)} if max_list is not None else None
        run(probe, f"{tmp}/dump.bin", "-i", f"{tmp}/q", "-d", f"{tmp}/db", "-o", f"{tmp}/out_y0.txt",
            *aln_args, env=env)
        if max_list is None:   # the CLI cannot express small budgets; keep y1/y2 for the others
            for y in (1, 2):
                run(ghostm, "aln", "-i", f"{tmp}/q", "-d", f"{tmp}/db", "-o", f"{tmp}/out_y{y}.txt",
                    "-y", str(y), *aln_args)
            run(ghostm, "aln", "-i", f"{tmp}/q", "-d", f"{tmp}/db", "-o", f"{tmp}/check.txt", *aln_args)
            assert open(f"{tmp}/check.txt", "rb").read() == open(f"{tmp}/out_y0.txt", "rb").read()
        db = formats.read_db(f"{tmp}/db")
        for i, ch in enumerate(db.chunks):   # pin my index builder while we are here
            kc, pos = formats.build_index(ch.seq, ch.seq_starts, ch.seed)
            assert (kc == ch.keys_count).all() and (pos == ch.positions).all(), "index mismatch"
            for ext in ("pos", "nam"):
                shutil.copy(f"{tmp}/db_{i}.{ext}", f"{out}/db_{i}.{ext}")
            with open(f"{tmp}/db_{i}.seq", "rb") as fi, \
                    gzip.GzipFile(f"{out}/db_{i}.seq.gz", "wb", mtime=0) as fo:
                fo.write(fi.read())
        json.dump({"seed": db.seed, "chunks": len(db.chunks), "sum_residues": db.sum_residues,
                   "max_chunk_len": db.max_chunk_len}, open(f"{out}/db.meta.json", "w"))
        for f in os.listdir(tmp):
            if f == "q.inf" or (f.startswith("q_") and f.split(".")[-1] in ("inf", "seq", "nam")):
                shutil.copy(f"{tmp}/{f}", f"{out}/{f}")
        for f in ("out_y0.txt", "out_y1.txt", "out_y2.txt"):
            if os.path.exists(f"{tmp}/{f}"):
                shutil.copy(f"{tmp}/{f}", f"{out}/{f}")
        with open(f"{tmp}/dump.bin", "rb") as fi, gzip.GzipFile(f"{out}/dump.bin.gz", "wb", mtime=0) as fo:
            fo.write(fi.read())
    json.dump({"aln_args": list(aln_args), "max_list_length": max_list, "l1024": l1024},
              open(f"{out}/meta.json", "w"))
    size = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
    print(f"{name}: {size / 1024:.0f} KiB")


def fasta(tmp, name, names, seqs, width=0):
    p = os.path.join(tmp, name)
    formats.write_fasta(p, names, seqs, width)
    return p


def main():
    # 1. the reference's own known answer (README.rdoc:138-149) and the literal config-1 reading
    make_case("readme_known_answer", f"{TESTSET}/queries.fasta", f"{TESTSET}/db.fasta",
              ["-t", "d"], [], [])
    make_case("testset_literal", f"{TESTSET}/db.fasta", f"{TESTSET}/queries.fasta", ["-t", "p"], [], [])
    with tempfile.TemporaryDirectory() as tmp:
        # 2. synthetic: 2 db chunks, runs of 6 equal names (translated-read stand-in), ragged lengths
        dbs, dbn = synth.protein_db(41, 90_000)
        qs, qn = synth.queries_from_db(42, dbs, 180, 60, group=6, min_length=25)
        make_case("synth_groups", fasta(tmp, "a.fa", dbn, dbs, 60), fasta(tmp, "aq.fa", qn, qs),
                  ["-l", "60", "-L", "1"], ["-l", "1"], [])
        # 2b. two db chunks: Merge carries hit lists from chunk 0 into chunk 1
        dbs2, dbn2 = synth.protein_db(43, 1_080_000)
        qs2, qn2 = synth.queries_from_db(44, dbs2, 150, 75, group=3)
        make_case("synth_two_chunks", fasta(tmp, "b.fa", dbn2, dbs2, 60), fasta(tmp, "bq.fa", qn2, qs2),
                  [], ["-l", "1"], ["-b", "12"])
        # 3. non-default options incl. best > 16 (introsort on carried lists), threshold 3
        make_case("synth_options", fasta(tmp, "a.fa", dbn, dbs, 60), fasta(tmp, "aq.fa", qn, qs),
                  ["-l", "60"], ["-l", "1"], ["-b", "24", "-r", "8", "-s", "1", "-e", "5", "-t", "3"])
        # 4. low-complexity db: large intervals, many candidates, candidate-chunk cuts, dropped tail
        dbs, dbn = synth.repeat_db(5, 60_000)
        qs, qn = synth.repeat_queries(6, 40, 75)
        make_case("repeats_chunked", fasta(tmp, "r.fa", dbn, dbs), fasta(tmp, "rq.fa", qn, qs), [], [],
                  [], max_list=700)
        # 5. long queries (config 4): needs the cap-raised reference build
        dbs, dbn = synth.protein_db(3, 120_000)
        qs, qn = synth.queries_from_db(4, dbs, 12, 400, min_length=150)
        make_case("long_queries_l1024", fasta(tmp, "l.fa", dbn, dbs, 60), fasta(tmp, "lq.fa", qn, qs),
                  ["-l", "400"], [], [], l1024=True)
        # 6. config 4's upper end: ragged queries of 300..1000 residues X-padded to 1000 (`qry -l 1000`,
        #    13 SW strips of 80 rows, list_len 499), two db chunks
        dbs, dbn = synth.protein_db(7, 1_300_000)
        qs, qn = synth.queries_from_db(8, dbs, 6, 1000, min_length=300)
        make_case("long_queries_l1000", fasta(tmp, "m.fa", dbn, dbs, 60), fasta(tmp, "mq.fa", qn, qs),
                  ["-l", "1000"], ["-l", "1"], [], l1024=True)


if __name__ == "__main__":
    main()
