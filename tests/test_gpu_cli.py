"""The `morsi` host program end to end on the GPU box: same bytes on stdout /
in the output file as the reference CLI produced (tests/golden/cli.json)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests.golden.make_golden import kat_input

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "imscript_b200", "lib", "morsi")


def test_cli_outputs_match_reference_bytes(golden_dir, tmp_path):
    gold = json.load(open(os.path.join(golden_dir, "cli.json")))
    fin = str(tmp_path / "in.npy")
    rgb = np.stack([kat_input(), kat_input()[::-1], kat_input()[:, ::-1]], axis=-1)
    np.save(fin, rgb)
    ran = 0
    for key, g in gold.items():
        argv = key.split() if key else []
        if g["rc"] != 0 or "IN" not in argv:
            continue
        argv = [fin if a == "IN" else a for a in argv]
        p = subprocess.run([CLI] + argv, capture_output=True)           # stdout form
        assert p.returncode == 0, (key, p.stderr)
        assert len(p.stdout) == g["stdout_len"], key
        assert hashlib.sha256(p.stdout).hexdigest()[:16] == g["stdout_sha"], key
        fout = str(tmp_path / "out.npy")                                  # file form
        p2 = subprocess.run([CLI] + argv + [fout], capture_output=True)
        assert p2.returncode == 0 and open(fout, "rb").read() == p.stdout, key
        p3 = subprocess.run([CLI] + argv[:2], input=open(fin, "rb").read(), capture_output=True)  # pipe form
        assert p3.returncode == 0 and p3.stdout == p.stdout, key
        ran += 1
    assert ran >= 4


def test_cli_multi_device_env_is_bit_identical(tmp_path):
    """MORSI_CUDA_DEVICES > visible devices is clamped; chunk height forced small
    so that several row-band chunks with halos are exercised."""
    fin = str(tmp_path / "in.npy")
    np.save(fin, kat_input())
    base = subprocess.run([CLI, "disk5", "tophat", fin], capture_output=True).stdout
    env = dict(os.environ, MORSI_CUDA_DEVICES="8", MORSI_CUDA_CHUNK_ROWS="7")
    again = subprocess.run([CLI, "disk5", "tophat", fin], capture_output=True, env=env).stdout
    assert base and base == again
