"""-m gpu: the drop-in boundary end to end.
  * ghostm_b200_aln (C++ host driver over the extended C ABI) must write the same output files as
    the reference's `ghostm aln`, for all three output styles, on one device and - when two are
    visible - with the db chunks spread over two devices;
  * oracle/_ref/ghostm_dropin = the UNMODIFIED reference host objects linked against
    libghostm_b200.so: `aln -D 0` drives the ten legacy symbols exactly as the reference does."""
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
                                  "synth_two_chunks", "synth_options"])
def test_host_driver_output_files(name, tmp_path):
    assert os.path.exists(ALN), "build with make -C ghostm_b200/csrc"
    meta, texts = _materialise(name, tmp_path)
    devices = ["0"] + (["0,1"] if _gpu_count() >= 2 else [])
    for dev in devices:
        for y, expect in texts.items():
            out = tmp_path / f"out_{y}_{dev.replace(',', '_')}.txt"
            subprocess.check_call([ALN, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o",
                                   str(out), "-D", dev, "-y", str(y)] + meta["aln_args"])
            assert open(out, encoding="latin-1", newline="").read() == expect, (name, dev, y)


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/ghostm_dropin not built")
@pytest.mark.parametrize("name", ["readme_known_answer", "synth_groups", "synth_two_chunks"])
def test_reference_host_linked_against_our_library(name, tmp_path):
    meta, texts = _materialise(name, tmp_path)
    out = tmp_path / "out.txt"
    # -l 1: the reference allocates 3 vectors of max_list_length per call in GPU mode
    subprocess.check_call([DROPIN, "aln", "-i", str(tmp_path / "q"), "-d", str(tmp_path / "db"), "-o", str(out),
                           "-D", "0", "-l", "1"] + meta["aln_args"])
    assert open(out, encoding="latin-1", newline="").read() == texts[0]
