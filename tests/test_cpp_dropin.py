"""The header-only C++ drop-in classes (include/hpxfft/...) exercised by ports of the reference's own
Catch2 tests (tests/cpp/*.cpp) and by the example drivers."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, gpu_count

CPP = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def cpp_bins():
    subprocess.run(["make", "-s", "-C", CPP], check=True)
    return CPP


def run(cmd, env=None, cwd=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, env=e, cwd=cwd, capture_output=True, text=True, timeout=300)


def test_cpp_headers_compile_and_vector_2d(cpp_bins):
    r = run([os.path.join(cpp_bins, "test_vector_2d")])
    assert r.returncode == 0 and "test_vector_2d ok" in r.stdout, r.stderr


@pytest.mark.parametrize("world", [1, 3])
def test_cpp_bootstrap_rendezvous(cpp_bins, world, tmp_path):
    """world_size-3 exchange through the shared-directory rendezvous (no HPX, no GPU)."""
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_PORT="4712",
                   HPXFFT_B200_RENDEZVOUS=str(tmp_path))
        procs.append(subprocess.Popen([os.path.join(cpp_bins, "test_bootstrap")], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "test_bootstrap ok" in o, o


def test_cpp_bootstrap_ignores_stale_files(cpp_bins, tmp_path):
    """A second run in the SAME directory with the SAME MASTER_PORT (and staggered starts) must not pick up the
    files the first run left behind: the session nonce is drawn afresh and acknowledged by every rank."""
    import time
    world = 3
    for attempt in range(2):
        procs = []
        for rank in (2, 1, 0) if attempt else (0, 1, 2):
            env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_PORT="4713",
                       HPXFFT_B200_RENDEZVOUS=str(tmp_path))
            procs.append(subprocess.Popen([os.path.join(cpp_bins, "test_bootstrap")], env=env, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True))
            time.sleep(0.05)
        for p in procs:
            o = p.communicate(timeout=120)[0]
            assert p.returncode == 0 and "test_bootstrap ok" in o, o
    # data files are removed after the closing barrier; only tiny session / barrier files remain
    left = [f for f in os.listdir(tmp_path) if "_uid_" in f or "_ipc_" in f]
    assert all("done" in f for f in left), left


@pytest.mark.gpu
def test_cpp_shared_loop_golden(cpp_bins):
    r = run([os.path.join(cpp_bins, "test_shared_loop")])
    assert r.returncode == 0 and "test_shared_loop ok" in r.stdout, r.stdout + r.stderr


def test_cpp_hpx_branches_compile_against_the_api_stub(cpp_bins):
    """-DHPXFFT_B200_WITH_HPX against tests/cpp/hpx_stub (the subset of the HPX API the headers use; HPX itself is absent here):
    vector_2d serialisation round trip and the HPX bootstrap calls.  CPU only."""
    r = run([os.path.join(cpp_bins, "test_hpx_branch")])
    assert r.returncode == 0 and "test_hpx_branch ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_hpx_branch_agas_futures(cpp_bins):
    r = run([os.path.join(cpp_bins, "test_hpx_branch"), "gpu"], env={"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0 and "test_hpx_branch ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_agas_clients(cpp_bins):
    r = run([os.path.join(cpp_bins, "test_agas")], env={"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0 and "test_agas ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("comm", ["scatter", "all_to_all", "p2p"])
def test_cpp_distributed_loop_one_locality(cpp_bins, comm):
    r = run([os.path.join(cpp_bins, "test_distributed_loop"), comm], env={"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0 and "test_distributed_loop ok" in r.stdout, r.stdout + r.stderr
    if comm == "scatter":
        assert "Specify communication scheme: scatter or all_to_all" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("comm", ["scatter", "all_to_all", "p2p"])
def test_cpp_distributed_loop_two_localities(cpp_bins, comm, tmp_path):
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_PORT="4711",
                   HPXFFT_B200_RENDEZVOUS=str(tmp_path))
        procs.append(subprocess.Popen([os.path.join(cpp_bins, "test_distributed_loop"), comm], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "test_distributed_loop ok" in o, o


@pytest.mark.gpu
def test_example_cli_csv_schema(cpp_bins, oracle, tmp_path):
    r = run([os.path.join(cpp_bins, "hpxfft_shared_loop"), "--nx=64", "--ny=128", "--header=1", "--result=1"], cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = (tmp_path / "runtimes" / "runtimes_hpx_shared_loop.txt").read_text().splitlines()
    # examples/hpxfft/shared_loop_2d.cpp:100-111
    assert lines[0] == ("n_threads;n_x;n_y;plan;run_flag;total;initialization;fft_2d_total;first_fftw;first_trans;"
                        "second_fftw;second_trans;plan_time;plan_flops;")
    f = lines[1].split(";")
    assert f[1:5] == ["64", "128", "estimate", "par"] and len(f) == 15
    assert "FFTW r2c 1D plan:" in (tmp_path / "plans" / "plan_hpx_shared_loop.txt").read_text()
    # the reference's default size (8 x 14) runs too
    r2 = run([os.path.join(cpp_bins, "hpxfft_shared_loop"), "--result=1"], cwd=tmp_path)
    assert r2.returncode == 0 and "(728 0) (-56 245.352)" in r2.stdout, r2.stdout + r2.stderr
    # --result prints the spectrum as "(re im)" pairs: row 0 of the ramp against the closed form
    row0 = [ln for ln in r.stdout.splitlines() if ln.startswith("(")][0]
    vals = np.array([float(t) for t in row0.replace("(", " ").replace(")", " ").split()])
    ref = np.asarray(oracle.ramp_analytic(64, 128)[0], dtype=np.float64)
    assert np.allclose(vals, ref, rtol=1e-5, atol=1e-3)
