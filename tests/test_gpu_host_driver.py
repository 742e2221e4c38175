"""-m gpu: the drop-in boundary end to end.
  * ghostm_b200_aln (C++ host driver over the extended C ABI) must write the same output files as
    the reference's `ghostm aln`, for all three output styles, on one device, with the db chunks
    and query slices spread over several shards of one device and - when two are visible - over
    two devices;
  * oracle/_ref/ghostm_dropin = the UNMODIFIED reference host objects linked against
    libghostm_b200.so: `aln -D 0` drives the ten legacy symbols exactly as the reference does;
  * oracle/_ref/ghostm_fast = the reference host with ONLY Aligner::Execute replaced by
    integration/aligner_b200.cpp (gm_* API): readers, SetOption, Statistics and WriteOutput* are the
    reference's own objects."""
import os
import shutil
import subprocess

import pytest

from ghostm_b200 import formats
from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALN = os.path.join(ROOT, "ghostm_b200", "ghostm_b200_aln")
DROPIN = os.path.join(ROOT, "oracle", "_ref", "ghostm_dropin")
FAST = os.path.join(ROOT, "oracle", "_ref", "ghostm_fast")


def _materialise(name, tmp_path):
    db, qchunks, kw, _, _, texts = H.golden(name)
    formats.write_db(str(tmp_path / "db"), db)
    d = os.path.join(H.GOLDEN_DIR, name)
    for f in os.listdir(d):
        if f == "q.inf" or f.startswith("q_"):
            shutil.copy(os.path.join(d, f), tmp_path / f)
    import json
    meta = json.load(open(os.path.join(d, "meta.json")))
    return meta, texts


def _gpu_count():
    from ghostm_b200 import capi
    return capi.load().gm_device_count()


@pytest.mark.parametrize("name", ["readme_known_answer", "testset_literal", "synth_groups",
                                  "synth_two_chunks", "synth_options", "long_queries_l1024",
                                  "long_queries_l1000"])
def test_host_driver_output_files(name, tmp_path):
    assert os.path.exists(ALN), "build with make -C ghostm_b200/csrc"
    meta, texts = _materialise(name, tmp_path)
    # "0,0": two shards on one device - the multi-device code path (front/back split, candidate
    # transfer by query slice) runs even when a single GPU is visible
    devices = ["0", "0,0", "0,0,0"] + (["0,1"] if _gpu_count() >= 2 else [])
    for dev in devices:
        for y, expect in texts.items():
            out = tmp_path / f"out_{y}_{dev.replace(',', '_')}.txt"
            subprocess.check_call([ALN, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o",
                                   str(out), "-D", dev, "-y", str(y)] + meta["aln_args"])
            assert open(out, encoding="latin-1", newline="").read() == expect, (name, dev, y)


def test_host_driver_streams_chunks_that_do_not_fit(tmp_path):
    """One device, db chunks not kept resident (forced): read, upload, align, release per query
    chunk like the reference (aligner.cpp:115-173) - same output bytes; -v reports the phases."""
    meta, texts = _materialise("synth_two_chunks", tmp_path)
    env = dict(os.environ, GHOSTM_B200_STREAM="1")
    for y, expect in texts.items():
        out = tmp_path / f"stream_{y}.txt"
        r = subprocess.run([ALN, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o", str(out),
                            "-D", "0", "-y", str(y), "-v"] + meta["aln_args"], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "(streamed)" in r.stdout and "Calculate scores ..." in r.stdout
        assert open(out, encoding="latin-1", newline="").read() == expect, y


def test_host_driver_device_error_is_not_a_finished_run(tmp_path):
    """A device-side failure gives a non-zero exit status and leaves no partial output file."""
    meta, _ = _materialise("synth_groups", tmp_path)
    out = tmp_path / "out.txt"
    r = subprocess.run([ALN, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o", str(out),
                        "-D", "99"], capture_output=True, text=True)
    assert r.returncode != 0 and "error" in r.stderr
    assert not out.exists()


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/ghostm_dropin not built")
@pytest.mark.parametrize("name", ["readme_known_answer", "synth_groups", "synth_two_chunks"])
def test_reference_host_linked_against_our_library(name, tmp_path):
    meta, texts = _materialise(name, tmp_path)
    out = tmp_path / "out.txt"
    # -l 1: the reference allocates 3 vectors of max_list_length per call in GPU mode
    subprocess.check_call([DROPIN, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o", str(out),
                           "-D", "0", "-l", "1"] + meta["aln_args"])
    assert open(out, encoding="latin-1", newline="").read() == texts[0]


@pytest.mark.skipif(not os.path.exists(FAST), reason="oracle/_ref/ghostm_fast not built")
@pytest.mark.parametrize("name", ["readme_known_answer", "testset_literal", "synth_groups",
                                  "synth_two_chunks", "synth_options", "long_queries_l1024",
                                  "long_queries_l1000"])
def test_reference_host_with_execute_on_gm_api(name, tmp_path):
    """The reference binary with Aligner::Execute swapped for integration/aligner_b200.cpp: same
    output bytes as the golden reference text for -y 0/1/2; -v prints the per-phase times."""
    meta, texts = _materialise(name, tmp_path)
    for y, expect in texts.items():
        out = tmp_path / f"fast_{y}.txt"
        r = subprocess.run([FAST, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o", str(out),
                            "-D", "0", "-y", str(y), "-v"] + meta["aln_args"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(out, encoding="latin-1", newline="").read() == expect, (name, y)
        assert "Calculate scores ..." in r.stdout and "Merge results ..." in r.stdout and "Complete." in r.stdout


REF = os.path.join(ROOT, "oracle", "_ref", "ghostm")


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ghostm")),
                    reason="oracle/_ref/ghostm not built (needs /root/reference)")
@pytest.mark.parametrize("k", [4, 3])
def test_db_command_writes_the_reference_files(k, tmp_path):
    """`ghostm_b200_aln db` (FASTA parsing + chunking on the host, key extraction + hand-written
    stable counting sort + CSR boundaries on the device) against the reference's `ghostm db`
    (db_creator.cpp:369-479) on a multi-chunk FASTA with the reader's corner cases: every output
    file byte-identical."""
    import numpy as np
    from ghostm_b200 import synth
    rng = np.random.default_rng(77)
    dbs, dbn = synth.protein_db(78, 2_400_000)
    letters = "ARNDCQEGHILKMFPSTWYVBJZX*"
    with open(tmp_path / "db.fa", "w", newline="") as f:
        f.write("junk before the first header\n")
        for i, (s, nm) in enumerate(zip(dbs, dbn)):
            txt = "".join(letters[int(c)] for c in s)
            if i % 7 == 0:
                txt = txt.lower()
            if i % 11 == 0:
                txt = txt[:20] + "XX-U" + txt[20:]          # X, and characters that map to X
            if i % 13 == 0:
                txt = txt[: int(rng.integers(1, 6))]       # not longer than the seed: not indexed
            eol = "\r\n" if i % 5 == 0 else "\n"
            head = (">  " if i % 3 == 0 else ">") + nm + (" some description" if i % 4 == 0 else "")
            f.write(head + eol)
            for a in range(0, len(txt), 60):
                f.write(txt[a:a + 60] + ("+" if i % 17 == 0 else "") + eol)
            if i % 19 == 0:
                f.write("\n")
    quiet = dict(stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call([REF, "db", "-i", str(tmp_path / "db.fa"), "-o", str(tmp_path / "ref"), "-l", "1",
                           "-k", str(k)], **quiet)
    subprocess.check_call([ALN, "db", "-i", str(tmp_path / "db.fa"), "-o", str(tmp_path / "ours"), "-l", "1",
                           "-k", str(k)], **quiet)
    ref_files = sorted(f for f in os.listdir(tmp_path) if f.startswith("ref"))
    assert len(ref_files) >= 1 + 5 * 3, ref_files
    for f in ref_files:
        ours = tmp_path / ("ours" + f[3:])
        assert ours.exists(), f
        assert ours.read_bytes() == (tmp_path / f).read_bytes(), f


_CODON = ["GCT", "CGT", "AAT", "GAT", "TGT", "CAA", "GAA", "GGT", "CAT", "ATT", "CTT", "AAA", "ATG",
          "TTT", "CCT", "TCT", "ACT", "TGG", "TAT", "GTT"]   # one codon per residue code A..V


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ghostm not built (needs /root/reference)")
@pytest.mark.parametrize("n_reads,style", [(100_000, "0"), (20_000, "2")])
def test_config2_standin_against_the_live_reference(n_reads, style, tmp_path):
    """BASELINE config 2 stand-in (testset/large_queries.fasta is absent): 75-nt DNA
    reads, half back-translated from db proteins with 10 % substitutions, half random, formatted
    by the REFERENCE's own `db` and `qry -t d` (6-frame translation, six same-name queries per
    read), aligned by the reference CPU aligner and by ghostm_b200_aln on the same files: the
    output files must be byte-identical."""
    import numpy as np
    from ghostm_b200 import synth
    rng = np.random.default_rng(20261018)
    dbs, dbn = synth.protein_db(21, 1_000_000)
    formats.write_fasta(str(tmp_path / "db.fa"), dbn, dbs, 70)
    concat = np.concatenate(dbs)
    from_db = rng.random(n_reads) < 0.5
    starts = rng.integers(0, concat.shape[0] - 25, size=n_reads)
    rand_nt = rng.integers(0, 4, size=(n_reads, 75))
    sub = rng.random((n_reads, 75)) < 0.10
    with open(tmp_path / "reads.fa", "w") as f:
        for i in range(n_reads):
            if from_db[i]:
                nt = list("".join(_CODON[min(int(c), 19)] for c in concat[starts[i]:starts[i] + 25]))
                for k in np.flatnonzero(sub[i]):
                    nt[k] = "ACGT"[rand_nt[i, k]]
                nt = "".join(nt)
            else:
                nt = "".join("ACGT"[x] for x in rand_nt[i])
            f.write(f">r{i}\n{nt}\n")
    quiet = dict(stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call([REF, "db", "-i", str(tmp_path / "db.fa"), "-o", str(tmp_path / "db")], **quiet)
    subprocess.check_call([REF, "qry", "-t", "d", "-i", str(tmp_path / "reads.fa"), "-o", str(tmp_path / "q")], **quiet)
    subprocess.check_call([REF, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o",
                           str(tmp_path / "ref.txt"), "-y", style], **quiet)
    ref = (tmp_path / "ref.txt").read_bytes()
    assert ref.count(b"\n") > n_reads // 4
    for dev in (["0", "0,0"] if n_reads <= 20_000 else ["0"]):
        subprocess.check_call([ALN, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o",
                               str(tmp_path / "ours.txt"), "-D", dev, "-y", style], **quiet)
        assert (tmp_path / "ours.txt").read_bytes() == ref, dev
    if os.path.exists(FAST):     # the reference host with Execute on the gm_* API, same files
        subprocess.check_call([FAST, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o",
                               str(tmp_path / "fast.txt"), "-D", "0", "-y", style], **quiet)
        assert (tmp_path / "fast.txt").read_bytes() == ref
