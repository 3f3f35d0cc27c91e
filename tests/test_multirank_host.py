"""CPU tests of the multi-rank host logic: canonical exchange plans (in-process, all ranks) and the NCCL driver's
message pattern replayed over a world_size-2 gloo group with CPU tensors."""
import os
import subprocess
import sys

from lambdapic_b200.multigpu import build_plan
from lambdapic_b200.workloads import make_patch_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plans_are_mutually_consistent_2d_and_3d():
    for dim, shape in ((3, (4, 2, 2)), (2, (4, 4, 1))):
        for nranks in (2, 4):
            grids = [make_patch_grid(dim, *shape, 8, 8, 8, 1.0, 1.0, 1.0, rank=r, nranks=nranks) for r in range(nranks)]
            plans = [build_plan(g) for g in grids]
            for a in range(nranks):
                peers, send, recv = plans[a]
                assert a not in peers
                for b in peers:
                    _, sb, rb = plans[b]
                    assert len(send[b]) == len(rb[a]) and len(recv[b]) == len(sb[a])
                    for (p, bd), (q, bq) in zip(send[b], rb[a]):
                        assert grids[a].neighbor_index[p, bd] == grids[b].index[q]
                        assert grids[b].neighbor_index[q, bq] == grids[a].index[p]
            for a in range(nranks):
                n_remote = int((grids[a].neighbor_rank >= 0).sum())
                assert n_remote == sum(len(v) for v in plans[a][2].values())


def test_nccl_driver_message_pattern_over_gloo(tmp_path):
    """drive_nccl with a toy program over gloo/CPU tensors: two exchanges, asymmetric sizes, one empty message."""
    script = tmp_path / "w.py"
    script.write_text(
        "import sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch, torch.distributed as dist\n"
        "import lambdapic_b200.multigpu as mg\n"
        "torch.cuda.synchronize = lambda *a, **k: None\n"
        "dist.init_process_group('gloo')\n"
        "r = dist.get_rank(); o = 1 - r\n"
        "def prog():\n"
        "    s = torch.arange(4 + r, dtype=torch.float64) + 10 * r\n"
        "    got = yield ('f64', {o: (s, 4 + r)}, {o: (torch.zeros(8, dtype=torch.float64), 4 + o)})\n"
        "    assert torch.equal(got[o][:4 + o], torch.arange(4 + o, dtype=torch.float64) + 10 * o)\n"
        "    got = yield ('f64', {o: (torch.zeros(1, dtype=torch.float64), 0)}, {o: (torch.zeros(1, dtype=torch.float64), 0)})\n"
        "    return 'fin'\n"
        "assert mg.drive_nccl(prog(), r) == 'fin'\n"
        "dist.barrier(); dist.destroy_process_group(); print('ok', r)\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
