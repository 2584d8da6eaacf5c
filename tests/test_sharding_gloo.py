"""Host-side logic of the N>1 path on CPU: two ranks (gloo) each generate their shard of the seeded
problem sequence and reduce the timing / count scalars the way bench.py does.  No GPU, no collective
on the data path -- the only cross-rank traffic is the final reduction."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from socp_b200 import generators as gen
from socp_b200 import sharding as sh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_tile_the_batch():
    for batch, world in [(10, 1), (10, 3), (1000, 8), (7, 8), (100_000, 8)]:
        plan = sh.gather_plan(batch, world)
        assert plan[0][0] == 0 and plan[-1][1] == batch
        for (a0, a1), (b0, b1) in zip(plan, plan[1:]):
            assert a1 == b0 and a0 <= a1
        assert sum(b - a for a, b in plan) == batch


def test_problem_bytes_do_not_depend_on_the_shard():
    full = gen.make_config("C3", batch=12)
    a = gen.make_config("C3", batch=5, first=0)
    b = gen.make_config("C3", batch=7, first=5)
    assert np.array_equal(np.concatenate([a.G_cm, b.G_cm]), full.G_cm)
    assert np.array_equal(np.concatenate([a.c, b.c]), full.c)
    assert np.array_equal(np.concatenate([a.h, b.h]), full.h)


WORKER = textwrap.dedent("""
    import os, sys, zlib
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "socp.jl_b200"))
    import numpy as np, torch, torch.distributed as dist
    from socp_b200 import generators as gen, sharding as sh
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    B = 6
    first, last = sh.weak_shard(B, rank)
    prob = gen.make_config("C2", batch=B, first=first)
    crc = zlib.crc32(prob.G_cm.tobytes()) ^ zlib.crc32(prob.c.tobytes())
    # what bench.py reduces: MAX of the timed regions, SUM of the counts
    tt = torch.tensor([0.5 + rank, 1.0, 2.0 - rank], dtype=torch.float64)
    cnt = torch.tensor([float(B), float(B - rank)], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    crcs = [None] * world
    dist.all_gather_object(crcs, (first, last, crc))
    if rank == 0:
        print("RESULT", (tt.tolist(), cnt.tolist(), crcs))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    tt, cnt, crcs = eval(line[len("RESULT"):])
    assert tt == [1.5, 1.0, 2.0] and cnt == [12.0, 11.0]
    # the two shards are the two halves of the single-process batch
    import zlib
    full = gen.make_config("C2", batch=12)
    for first, last, crc in crcs:
        want = zlib.crc32(full.G_cm[first:last].tobytes()) ^ zlib.crc32(full.c[first:last].tobytes())
        assert crc == want
    assert [c[:2] for c in crcs] == [(0, 6), (6, 12)]


def test_numa_binding_helpers(tmp_path):
    """bind_host_to_pci_device on a fake sysfs tree: cpulist parsing, unknown devices / nodes leave the process alone."""
    import os
    from socp_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:3b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices" / "system" / "node" / "node1"
    node.mkdir(parents=True)
    mine = sorted(os.sched_getaffinity(0))
    (node / "cpulist").write_text(",".join(str(c) for c in mine) + "\n")
    assert sharding.numa_node_of_pci_device("0000:3B:00.0", str(tmp_path)) == 1
    assert sharding.numa_node_of_pci_device("0000:00:00.0", str(tmp_path)) == -1
    before = os.sched_getaffinity(0)
    assert sharding.bind_host_to_pci_device("0000:3b:00.0", str(tmp_path)) == 1
    assert os.sched_getaffinity(0) == before                    # the node's CPUs are exactly the allowed ones here
    (dev / "numa_node").write_text("-1\n")
    assert sharding.bind_host_to_pci_device("0000:3b:00.0", str(tmp_path)) == -1
    (dev / "numa_node").write_text("1\n")
    (node / "cpulist").write_text("100000\n")                   # no overlap with the allowed CPUs: nothing changes
    assert sharding.bind_host_to_pci_device("0000:3b:00.0", str(tmp_path)) == -1
    assert os.sched_getaffinity(0) == before
