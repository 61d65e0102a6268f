"""Tensor-parallel host logic on CPU: shard plans partition every tensor exactly, and a
world_size-2 gloo run of the emulated TP forward reproduces the full oracle forward."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,size", [("tiny-untied", 2), ("8b", 2), ("8b", 4), ("8b", 8), ("32b", 8), ("4b", 8)])
def test_shard_plan_partitions_exactly(pkg, name, size):
    sh = pkg.checkpoint.SHAPES[name]
    shape = dict(n_heads=sh.n_heads, n_kv_heads=sh.n_kv_heads, head_dim=sh.head_dim, hidden_dim=sh.hidden_dim,
                 vocab_size=sh.vocab_size)
    plans = [pkg.tp.shard_plan(shape, r, size) for r in range(size)]
    for field, total in (("q_rows", sh.proj_dim), ("kv_rows", sh.kv_dim), ("hid_rows", sh.hidden_dim),
                         ("o_cols", sh.proj_dim), ("vocab_rows", sh.vocab_size), ("kv_heads", sh.n_kv_heads)):
        covered = np.concatenate([np.arange(getattr(p, field).start, getattr(p, field).stop) for p in plans])
        assert np.array_equal(covered, np.arange(total)), field
    for p in plans:  # row-parallel windows keep whole Q8_0 groups, query heads stay with their kv head
        assert p.o_cols.start % 64 == 0 and len(p.o_cols) % 64 == 0
        assert p.hid_rows.start % 64 == 0 and len(p.hid_rows) % 64 == 0
        kv_mul = sh.n_heads // sh.n_kv_heads
        assert p.q_rows.start // sh.head_dim == p.kv_heads.start * kv_mul


def test_shard_plan_rejects_indivisible(pkg):
    with pytest.raises(ValueError):
        pkg.tp.shard_plan(dict(n_heads=16, n_kv_heads=8, head_dim=128, hidden_dim=3072, vocab_size=151936), 0, 16)


def test_load_views_matches_oracle_reader(pkg, oracle, ckpt_dir):
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "tiny-untied", seed=11)
    v = pkg.checkpoint.load_views(path)
    with oracle.open(path, 8) as om:
        assert v["dim"] == om.p.dim and v["n_layers"] == om.p.n_layers
        n = om.p.dim
        got = np.ctypeslib.as_array(om.p.out_norm, shape=(n,))
        assert np.array_equal(got, v["out_norm"])
        q0 = np.ctypeslib.as_array(om.p.cls.q, shape=(om.p.vocab_size * n,))
        assert np.array_equal(q0, v["cls"][0].reshape(-1))


def test_emulated_tp2_forward_matches_oracle_gloo(pkg, ckpt_dir):
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "tiny-untied", seed=11)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tests", "tp_worker.py"), path],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0 and "TP_OK world=2" in out.stdout, out.stdout[-2000:]
