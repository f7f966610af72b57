"""CPU: host-side logic -- synthetic generator, algorithmic-byte accounting, frame sharding across ranks
(world_size 2 over gloo), bench reference arm."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gtsam_vslam_b200 import synth  # noqa: E402


def test_synth_is_deterministic_and_shifted():
    a = synth.stereo_pair(320, 240, 8, 3)
    b = synth.stereo_pair(320, 240, 8, 3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = synth.stereo_pair(320, 240, 8, 4)
    assert not np.array_equal(a[0], c[0])
    # right = canvas shifted by d: x_R = x_L - d (up to the independent noise)
    L, R = a
    d = np.abs(L[:, 8:].astype(int) - R[:, :-8].astype(int))
    assert np.median(d) <= 3
    seq = synth.shifted_sequence(160, 120, 3, seed=1)
    assert np.median(np.abs(seq[0][1:, 2:].astype(int) - seq[1][:-1, :-2].astype(int))) <= 3


def test_algorithmic_bytes_match_survey_table():
    # SURVEY.md section 8(d) table
    for name, A, b_img, b_frame in (("kitti", 1444097, 6837539, 13701078), ("euroc", 1117367, 5257735, 10528470),
                                    ("hires", 9739318, 45785474, 91674948)):
        cfg = synth.CONFIGS[name]
        st, bi, bf = bench.algorithmic_bytes(cfg)
        assert st["fast"] == A and bi == b_img and bf == b_frame


WORKER = r"""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import numpy as np
from gtsam_vslam_b200 import synth
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
F = 2
L, R = synth.stereo_batch(160, 120, 4, range(rank * F, rank * F + F))   # bench.make_frames sharding rule
# host gather of per-rank results only (no data-path collective): here the per-frame checksums
mine = torch.tensor([int(L[i].sum()) for i in range(F)], dtype=torch.int64)
out = [torch.zeros(F, dtype=torch.int64) for _ in range(world)]
dist.all_gather(out, mine)
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)   # timing = max over ranks
if rank == 0:
    print(json.dumps({"sums": torch.cat(out).tolist(), "tmax": float(t)}))
dist.destroy_process_group()
"""


def test_frame_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29531", str(script), ROOT], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    # rank r owns global frames [r*F, (r+1)*F): together the ranks cover the global batch exactly once, in order
    L, _ = synth.stereo_batch(160, 120, 4, range(4))
    assert res["sums"] == [int(L[i].sum()) for i in range(4)]
    assert res["tmax"] == 2.0


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0


def test_multi_device_sharding_rule_world_size_2_gloo(tmp_path):
    """gvs_multi / bench.py shard a batch into contiguous ranges [g*n/G, (g+1)*n/G); two gloo ranks that each take the range of
    their rank cover the batch exactly once and a host gather restores the batch order (the rule csrc/multi.cu implements)."""
    script = tmp_path / "w2.py"
    script.write_text(r'''
import json, sys
import torch, torch.distributed as dist
dist.init_process_group("gloo")
r, G = dist.get_rank(), dist.get_world_size()
n = 7
first = [g * n // G for g in range(G + 1)]
mine = torch.full((n,), -1, dtype=torch.int64)
mine[first[r]:first[r + 1]] = torch.arange(first[r], first[r + 1]) * 10 + r     # "results" of the rank's frames
out = [torch.zeros(n, dtype=torch.int64) for _ in range(G)]
dist.all_gather(out, mine)                                                       # host gather
merged = torch.stack(out).max(0).values
if r == 0:
    print(json.dumps({"merged": merged.tolist(), "first": first}))
dist.destroy_process_group()
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["first"] == [0, 3, 7]
    assert res["merged"] == [0, 10, 20, 31, 41, 51, 61]


def test_multi_entry_point_fails_loudly_without_a_device():
    import gtsam_vslam_b200 as gvs
    if gvs.device_count() > 0:
        return
    import pytest
    with pytest.raises(gvs.GvsError):
        gvs.MultiGpuTracker(400.0, 0.1, [0, 1], width=320, height=240, frames_per_device=2, nFeatures=300, nLevels=4)
