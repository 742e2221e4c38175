"""CPU: the oracle (oracle/ghostm_oracle.c) against the committed golden vectors made from the
unmodified reference, and - when oracle/_ref is present - against fresh reference runs."""
import os
import subprocess

import numpy as np
import pytest

from ghostm_b200 import formats, synth
from oracle import oracle as O
from tests import helpers as H


@pytest.mark.parametrize("name", H.GOLDEN_CASES)
def test_oracle_matches_golden(name):
    db, qchunks, kw, stages, results, texts = H.golden(name)
    opt = O.Options(**kw)
    si = 0
    text = ""
    for qi, qc in enumerate(qchunks):
        mine = []
        res = O.align_chunk(qc, db, opt, mine)
        for (ci, cc, ids, starts, scores, ends) in mine:
            r = stages[si]
            si += 1
            assert (r[0], r[1], r[2]) == (qi, ci, cc)
            assert np.array_equal(r[3], ids) and np.array_equal(r[4], starts)
            assert np.array_equal(r[5], scores) and np.array_equal(r[6], ends)
        ref = results[qi]
        for i in range(qc.n):
            got = res.hits[i, :res.counts[i]]
            assert len(ref[i]) == got.shape[0]
            for h, g in zip(ref[i], got):
                assert (h["db_id"], h["score"], h["db_start"], h["db_end"], h["aln_len"],
                        h["aln_match"]) == (g["db_id"], g["score"], g["db_start"], g["db_end"],
                                            g["aln_len"], g["aln_match"])
                assert np.float32(h["seq_id"]).tobytes() == np.float32(g["seq_id"]).tobytes()
                assert h["db_name"] == db.chunks[int(g["db_chunk"])].names[int(g["db_id"])]
        text += O.format_output(res, qc, db, opt)
        if 1 in texts and len(qchunks) == 1:
            assert H.format_v1(res.lists(), qc, db) == texts[1]
    assert si == len(stages)
    assert text == texts[0]


def test_readme_known_answer_numbers():
    """README.rdoc:138-149: the numeric columns of the reference's published sample output."""
    _, _, _, _, _, texts = H.golden("readme_known_answer")
    rows = [r.split("\t")[2:9] for r in texts[0].strip("\n").split("\n")]
    expect = [["100", "25", "25", "1", "25", "2.75456e-15", "60.4622"],
              ["100", "10", "10", "16", "25", "2.58417e-05", "27.335"],
              ["100", "24", "24", "1", "24", "1.36707e-14", "58.151"],
              ["100", "9", "9", "16", "24", "0.000128251", "25.0238"],
              ["100", "25", "25", "1", "25", "4.55093e-10", "43.1282"],
              ["84.2105", "19", "16", "1", "19", "1.15998e-05", "28.4906"],
              ["100", "25", "25", "1", "25", "2.85052e-12", "50.447"],
              ["84.2105", "19", "16", "7", "25", "1.15998e-05", "28.4906"],
              ["100", "10", "10", "16", "25", "2.58417e-05", "27.335"],
              ["100", "25", "25", "1", "25", "2.85052e-12", "50.447"],
              ["84.2105", "19", "16", "7", "25", "1.15998e-05", "28.4906"],
              ["100", "10", "10", "16", "25", "2.58417e-05", "27.335"]]
    assert rows == expect


def test_blosum62_tables_agree():
    from ghostm_b200 import workloads
    assert np.array_equal(O.blosum62(), workloads.blosum62())


@pytest.mark.skipif(O.ref_bin("ref_probe") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_fresh_reference_run(tmp_path):
    """Seeded synthetic input through the reference's own db/qry/aln, multi-chunk, with cuts."""
    dbs, dbn = synth.protein_db(71, 1_300_000)
    qs, qn = synth.queries_from_db(72, dbs, 240, 75, group=6)
    formats.write_db(str(tmp_path / "db"), formats.make_db(dbs, dbn, 4, 0.5))
    formats.write_queries(str(tmp_path / "q"), formats.make_query_chunks(qs, qn, 75, 0.006))
    env = dict(os.environ, GMPROBE_MAX_LIST_LENGTH="400")
    subprocess.check_call([O.ref_bin("ref_probe"), str(tmp_path / "dump.bin"), "-i", str(tmp_path / "q"),
                           "-d", str(tmp_path / "db"), "-o", str(tmp_path / "out.txt")], env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    stages, _ = O.read_probe_dump(str(tmp_path / "dump.bin"))
    db = formats.read_db(str(tmp_path / "db"))
    opt = O.Options(max_list_length=400)
    text, si = "", 0
    for qc in formats.read_queries(str(tmp_path / "q")):
        mine = []
        res = O.align_chunk(qc, db, opt, mine)
        for (_, _, ids, starts, scores, ends) in mine:
            r = stages[si]
            si += 1
            assert np.array_equal(r[3], ids) and np.array_equal(r[4], starts)
            assert np.array_equal(r[5], scores) and np.array_equal(r[6], ends)
        text += O.format_output(res, qc, db, opt)
    assert si == len(stages) and si > 3
    assert text == open(tmp_path / "out.txt", encoding="latin-1", newline="").read()
