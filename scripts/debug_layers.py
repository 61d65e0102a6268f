import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from oracle.binding import Oracle
pkg = load_pkg(); ql = pkg.QwenLib(); orc = Oracle()
shape = sys.argv[1] if len(sys.argv) > 1 else "tiny-untied"
npos = int(sys.argv[2]) if len(sys.argv) > 2 else 3
path = pkg.checkpoint.ensure_checkpoint("/tmp/qwen3_b200_ckpt", shape, seed=11)
sh = pkg.checkpoint.SHAPES[shape]
toks = np.random.default_rng(2).integers(0, sh.vocab_size, size=npos)
om = orc.open(path, 64, trace=True)
for pos, t in enumerate(toks): om.forward(int(t), pos)
tr = om.trace()
res = {}
for psel in (1, 0):
    for n in range(1, sh.n_layers + 1):
        with ql.open(path, 64) as gm:
            gm.set_path(psel)
            for pos, t in enumerate(toks[:-1]): gm.forward(int(t), pos)   # fill cache with all layers
            gm.set_layers(n)
            gm.forward(int(toks[-1]), npos - 1)
            res[(psel, n)] = dict(x=gm.debug_read("x", sh.dim), att=gm.debug_read("att", sh.proj_dim),
                                  h=gm.debug_read("h", sh.hidden_dim), qkv=gm.debug_read("qkv", sh.proj_dim + 2 * sh.kv_dim))
for n in range(1, sh.n_layers + 1):
    a, b = res[(1, n)], res[(0, n)]
    msg = f"layers={n}:"
    for k in ("qkv", "att", "h", "x"):
        d = np.abs(a[k] - b[k]); msg += f" {k} ops-vs-mega {d.max():.2e} (n>{1e-5:g}: {(d>1e-5).sum()})"
    msg += f" | x final-norm vs oracle: mega {np.abs(b['h'] - tr['h'][n-1]).max():.2e} ops {np.abs(a['h'] - tr['h'][n-1]).max():.2e}"
    print(msg)
    d = np.abs(a['att'] - b['att'])
    if d.max() > 1e-5: print("   att bad idx", np.nonzero(d > 1e-5)[0][:20], "heads", sorted(set((np.nonzero(d>1e-5)[0]//128).tolist())))
    d = np.abs(a['qkv'] - b['qkv'])
    if d.max() > 1e-5: print("   qkv bad idx", np.nonzero(d > 1e-5)[0][:20])
