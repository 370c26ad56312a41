"""CPU, world_size 2 over gloo: the N>1 path of bench.py -- replicas with distinct sequences, MAX-over-ranks timing,
SUM-over-ranks frame counts, reference arm only on rank 0 (no collective on the data path)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
ms = [10.0 + 5.0 * rank, 20.0 - 3.0 * rank]
cnt = [100.0, 90.0 + rank]
t, v = bench.aggregate_over_ranks(dist, ms, cnt, "cpu")
seeds = [None] * world
dist.all_gather_object(seeds, bench.sequence_seed(rank))
if rank == 0:
    print(json.dumps({"t": t, "v": v, "seeds": seeds}))
dist.destroy_process_group()
""" % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_aggregation(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    d = json.loads(outs[0][0].strip().splitlines()[-1])
    assert d["t"] == [15.0, 20.0]          # max over ranks
    assert d["v"] == [200.0, 181.0]        # sum over ranks
    assert d["seeds"] == [1300, 1301]      # distinct sequences per rank


def test_reference_arm_runs_only_on_rank0():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_describe_the_same_workload():
    """the `config` object is built by one function for the GPU arm and the reference arm (the driver compares them),
    names the workload of BASELINE.json's configs and says how L2 is taken out of the measurement"""
    sys.path.insert(0, ROOT)
    import bench
    for name, cfg in bench.CONFIGS.items():
        a = bench.config_dict(cfg, name, 11032)
        b = bench.config_dict(dict(cfg), name, 11032)
        assert a == b and a["name"] == name and a["width"] == cfg["w"] and a["height"] == cfg["h"]
        assert "larger than L2" in a["l2"] and str(cfg["ring"]) in a["l2"]
        assert 2 * 3 * cfg["w"] * cfg["h"] * cfg["ring"] > 126e6  # the ring of inputs does exceed the 126 MB L2
    assert "1280x1024" in bench.CONFIGS["B"]["workload"] and "2448x2048" in bench.CONFIGS["E"]["workload"]


def test_numa_binding_is_best_effort():
    """multi-GPU runs bind a rank to its GPU's NUMA node when sysfs shows one; anything unreadable means 'leave the
    affinity alone', never an error"""
    sys.path.insert(0, ROOT)
    import bench

    class NoTopology:
        class cuda:
            @staticmethod
            def get_device_properties(i):
                raise RuntimeError("no device")
    before = os.sched_getaffinity(0)
    assert bench.bind_to_gpu_numa_node(NoTopology, 0) is None
    assert os.sched_getaffinity(0) == before
