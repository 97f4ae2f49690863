"""INTEGRATION.md section 2, compiled and run: the UNMODIFIED reference driver (its own option parser, .tri reader,
BuildTree, InitRadiosity, PrePartitionTriangles and PrintOut; oracle/ref_harness/tu_rad_radk.cpp includes
/root/reference/src/rad.cpp textually) with only the DoRadiosity loop (rad.cpp:497-516) replaced by the radk_ctx_step
calls the guide shows, linked against the product's libradk.so.  Every file it writes - all snapshots and the .fin -
must be byte-identical to what the reference's own binary (oracle/_ref/rad_oracle, CPU loop) writes."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "oracle", "_ref", "rad_ref_radk")
REF = os.path.join(ROOT, "oracle", "_ref", "rad_oracle")


def run(binary, args, cwd):
    return subprocess.run([binary] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)


@pytest.mark.gpu
@pytest.mark.parametrize("name,extra", [("sbox", ["-write_cycle", "100"]), ("church", ["-write_cycle", "50"]),
                                       ("model1s", ["-write_cycle", "7", "-converge", "2000", "-light", "0.8"])])
def test_reference_host_with_radk_loop_writes_the_reference_files(models_dir, tmp_path, name, extra):
    if not (os.path.exists(HOST) and os.path.exists(REF)):
        pytest.skip("oracle/_ref binaries did not travel with the repository (build them where /root/reference exists)")
    mine, ref = tmp_path / "mine", tmp_path / "ref"
    mine.mkdir(), ref.mkdir()
    args = extra + [str(models_dir / f"{name}.tri"), "-o", "out"]
    r = run(HOST, args, mine)
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-2000:]
    q = run(REF, ["-jobs", str(min(16, os.cpu_count() or 1))] + args, ref)
    assert q.returncode == 0
    files = sorted(os.listdir(ref))
    assert "out.fin" in files and files == sorted(os.listdir(mine))
    for f in files:
        assert (mine / f).read_bytes() == (ref / f).read_bytes(), f
    it = lambda b: re.findall(rb"Complete in #Iterations (\d+)", b)
    assert it(r.stderr) == it(q.stderr) and it(r.stderr)


def test_reference_host_binary_links_the_product_library():
    """CPU tier: the binary exists where the reference tree was available and resolves libradk.so in-tree."""
    if not os.path.exists(HOST):
        pytest.skip("oracle/_ref/rad_ref_radk not built (no /root/reference on this machine)")
    out = subprocess.run(["ldd", HOST], stdout=subprocess.PIPE, text=True).stdout
    line = [l for l in out.splitlines() if "libradk.so" in l]
    assert line and "hw-radiosity_b200/libradk.so" in line[0] and "not found" not in line[0]
    syms = subprocess.run(["nm", "-D", "--undefined-only", HOST], stdout=subprocess.PIPE, text=True).stdout
    for s in ("radk_ctx_create", "radk_ctx_step", "radk_ctx_download", "radk_ctx_destroy"):
        assert s in syms
