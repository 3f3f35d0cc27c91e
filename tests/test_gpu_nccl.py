"""Real multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): torchrun + NCCL send/recv of the packed
staging buffers, N ranks against the 1-rank reference golden vectors."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_nccl_ranks_match_single_rank_reference(nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "nccl_worker.py")],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert out.stdout.count("nccl-parity ok") == nranks


@pytest.mark.parametrize("nranks", [2, 4])
def test_public_api_simulation_on_nccl_ranks(nranks):
    """Simulation3D + MultiRankMPI (the reference's sim.mpi interface over NCCL) against the 1-rank reference."""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "tests", "nccl_sim_worker.py")],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert out.stdout.count("nccl-sim-parity ok") == nranks


@pytest.mark.parametrize("nranks", [2])
def test_moving_window_on_nccl_ranks_matches_single_rank(nranks):
    """MovingWindow with several ranks (callback/utils.py:648-716): allgathered patch positions, exchange plan rebuilt after
    every shift (lpic_halo_plan + lpic_comm_update), device-side recycle -- N ranks equal one rank."""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
                          "--master-addr", "127.0.0.1", "--master-port", "29549", os.path.join(ROOT, "tests", "nccl_mw_worker.py")],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "nccl-mw-parity ok" in out.stdout
