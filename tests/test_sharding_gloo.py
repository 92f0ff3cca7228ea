"""N>1 path on CPU: two gloo ranks take disjoint lat-band slabs (no data-path collective), and the
only exchanges are the bench's barrier + max-over-ranks timing reduction."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_slab_partition_and_timing_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from xsdba_b200.sharding import lat_band
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        lo, hi = lat_band(721, r, w)
        owned = torch.zeros(721, dtype=torch.int64); owned[lo:hi] = 1
        dist.all_reduce(owned)                       # test-only check: every row owned exactly once
        assert int(owned.min()) == 1 and int(owned.max()) == 1
        t = torch.tensor([10.0 + r], dtype=torch.float64)
        dist.barrier(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == 10.0 + w - 1
        if r == 0: print(json.dumps({{"rows": [lo, hi], "max_ms": float(t)}}))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    assert '"rows": [0, 360]' in res.stdout


def test_lat_band_covers_everything():
    sys.path.insert(0, ROOT)
    from xsdba_b200.sharding import lat_band, slabs
    for n, w in ((721, 8), (721, 3), (10, 4), (5, 8)):
        rows = []
        for r in range(w):
            lo, hi = lat_band(n, r, w)
            rows += list(range(lo, hi))
        assert rows == list(range(n))
    assert slabs(100, 48) == [(0, 48), (48, 48), (96, 4)]
