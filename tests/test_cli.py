"""The `rad` command line is the drop-in surface (reference src/rad.cpp:525-661): same options, same
outputs.  GPU tests run the shipped binary and compare every file it writes with what the unmodified
reference binary writes (oracle/_ref/rad_oracle, when it travelled with the repo) and with the goldens."""
import hashlib
import os
import re
import subprocess

import pytest

from conftest import ROOT

RAD = os.path.join(ROOT, "hw-radiosity_b200", "rad")
REF = os.path.join(ROOT, "oracle", "_ref", "rad_oracle")


def run(binary, args, cwd):
    return subprocess.run([binary] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)


def test_usage_and_missing_file(tmp_path):
    assert os.path.exists(RAD), "build the product first (make -C hw-radiosity_b200)"
    r = run(RAD, [], tmp_path)
    assert r.returncode == 1 and b"Usage:" in r.stderr and b"-parallel_src <integer>" in r.stderr
    r = run(RAD, ["nonexistent.tri", "-o", "x"], tmp_path)
    assert r.returncode == 1 and b"File open error::" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,extra", [("model2s", []), ("sbox", ["-write_cycle", "100"]),
                                       ("model1s", ["-write_cycle", "7", "-converge", "2000", "-light", "0.8"]),
                                       ("sbox", ["-parallel_src", "4", "-write_cycle", "50"])])
def test_outputs_match_reference_binary(models_dir, golden, tmp_path, name, extra):
    mine, ref = tmp_path / "mine", tmp_path / "ref"
    mine.mkdir(), ref.mkdir()
    args = extra + [str(models_dir / f"{name}.tri"), "-o", "out"]
    r = run(RAD, args, mine)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    files = sorted(os.listdir(mine))
    assert "out.fin" in files
    if not extra:
        fin = hashlib.sha256((mine / "out.fin").read_bytes()).hexdigest()
        assert fin == golden["models"][name]["full"]["fin_sha"]
    if os.path.exists(REF):
        q = run(REF, ["-jobs", "4"] + args, ref)
        assert q.returncode == 0
        assert files == sorted(os.listdir(ref))
        for f in files:
            assert (mine / f).read_bytes() == (ref / f).read_bytes(), f
        # same report on stdout (config table) and the same iteration count on stderr
        strip = lambda b: re.sub(rb"-jobs\s+\d+", b"-jobs N", b)
        assert strip(r.stdout) == strip(q.stdout)
        it = lambda b: re.findall(rb"Complete in #Iterations (\d+)", b)
        assert it(r.stderr) == it(q.stderr) and it(r.stderr)
        loops = lambda b: re.findall(rb"Loop (\d+) BEGIN", b)
        assert loops(r.stderr) == loops(q.stderr)


@pytest.mark.gpu
def test_zip_option(models_dir, tmp_path):
    r = run(RAD, ["-zip", str(models_dir / "model2s.tri"), "-o", "z"], tmp_path)
    assert r.returncode == 0
    assert os.path.exists(tmp_path / "z.fin.gz") and b"Compress output file z.fin ." in r.stdout


@pytest.mark.gpu
def test_split_switch_matches_split_enabled_reference(models_dir, golden, tmp_path):
    """`rad -split` = the reference compiled without DONT_PARTITION/PRE_PARTITION; exact when no random sample is drawn."""
    g = golden["split_variants"]["blocks_split_nosample"]
    r = run(RAD, ["-split"] + g["argv"] + [str(models_dir / "blocks.tri"), "-o", "s"], tmp_path)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    assert re.findall(rb"Complete in #Iterations (\d+)", r.stderr) == [str(g["shots"] + 1).encode()]
    text = (tmp_path / "s.fin").read_bytes()
    assert text.count(b"Triangle \n") == g["leaves"]


@pytest.mark.gpu
def test_multi_gpu_switch(models_dir, golden, tmp_path):
    """`rad -gpus N`: receiver shards on N devices of the box, one host thread each; same .fin bytes."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r = run(RAD, ["-gpus", str(n), "-write_cycle", "100", str(models_dir / "sbox.tri"), "-o", "m"], tmp_path)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    assert hashlib.sha256((tmp_path / "m.fin").read_bytes()).hexdigest() == golden["models"]["sbox"]["full"]["fin_sha"]
    assert sorted(os.listdir(tmp_path)) == ["m.100", "m.200", "m.300", "m.400", "m.500", "m.fin"]
