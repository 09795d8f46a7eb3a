"""Host-side logic that needs no GPU: the synthetic workload generator, the algorithmic-byte model, bench.py's
reference arm, and the multi-rank plumbing of bench.py on gloo with world_size 2."""
import dataclasses
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dist", ["uniform", "blocky", "rare", "stripes"])
def test_numpy_and_torch_generators_agree(dist):
    import torch

    from superterrainplus_b200 import workloads

    wl = workloads.Workload("t", (20, 12), (3, 3), 4, 9, 3, dist)
    t = workloads.make_maps_torch(wl, 5, 3, "cpu")
    for i in range(3):
        a = workloads.make_map_np(wl, 5 + i)
        assert a.dtype == np.uint16 and a.shape == (36, 60)
        assert np.array_equal(a, t[i].numpy())
        assert a.max() < 9
    assert not np.array_equal(workloads.make_map_np(wl, 0), workloads.make_map_np(wl, 1)) or dist == "stripes"


def test_baseline_configs_and_algorithmic_bytes():
    from superterrainplus_b200 import workloads

    c3 = workloads.CONFIGS["C3"]
    assert (c3.map_size, c3.radius, c3.biomes, c3.chunks) == ((512, 512), 64, 64, 256)
    assert workloads.CONFIGS["C2"].total == (3072, 3072)
    # C1 uniform = 0.664 + 1.049 + 16.777 MB per chunk
    c1 = workloads.CONFIGS["C1"]
    assert workloads.algorithmic_bytes(c1, 8 * 512 * 512) == 2 * 576 * 576 + 4 * (512 * 512 + 1) + 8 * 8 * 512 * 512
    # C3 uniform: 136.1 MB per chunk
    per_chunk = workloads.algorithmic_bytes(dataclasses.replace(c3, chunks=1), 64 * 512 * 512)
    assert abs(per_chunk / 1e6 - 136.1) < 0.1


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                          "--steps", "1", "--warmup", "0", "--cpu-seconds", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "filtered_mpixels_per_s" and line["unit"] == "Mpixels/s"
    assert line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SHF_ROOT"])
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# the two pieces of bench.py's N>1 path that do not need a GPU: chunk ownership and the max-over-ranks reduction
first, count = bench.shard_of(rank, world, 256, "weak")
assert (first, count) == (rank * 256, 256), (first, count)
first, count = bench.shard_of(rank, world, 256, "strong")
assert (first, count) == (rank * 128, 128), (first, count)
t = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == 10.0 + world - 1
dist.barrier()
if rank == 0:
    print("gloo ok")
dist.destroy_process_group()
'''


def test_two_rank_plumbing_on_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, SHF_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gloo ok" in out.stdout


def test_l2_rule_is_decided_from_the_workload_alone():
    """`config.l2` must say what the GPU arm does and be the same string in both arms: the rule only looks at the
    workload's definition. The default line (C3 uniform) exceeds the L2 on 1 to 8 GPUs; single sparse chunks do not."""
    sys.path.insert(0, ROOT)
    import bench
    from superterrainplus_b200 import workloads

    c3 = workloads.CONFIGS["C3"]
    for n_gpus in (1, 2, 4, 8):
        cfg = bench.config_of(c3, n_gpus, "strong")
        assert cfg["chunks_per_gpu"] == 256 // n_gpus and cfg["chunks_total"] == 256
        assert bench.outputs_exceed_l2(c3, cfg["chunks_per_gpu"]) and "exceed" in cfg["l2"]
        assert bench.config_of(c3, n_gpus, "weak")["chunks_total"] == 256 * n_gpus
    c1 = workloads.CONFIGS["C1"]
    assert not bench.outputs_exceed_l2(c1, 1) and "rewritten between steps" in bench.config_of(c1, 1, "strong")["l2"]
    blocky = dataclasses.replace(workloads.CONFIGS["C2"], dist="blocky")
    assert not bench.outputs_exceed_l2(blocky, 1)
    # a lower bound: never claims more than the histograms really hold (uniform C3: 64 bins per pixel)
    assert 4 + 8 * 32 <= workloads.algorithmic_bytes(dataclasses.replace(c3, chunks=1), 64 * 512 * 512) / (512 * 512)
