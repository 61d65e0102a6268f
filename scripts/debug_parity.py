import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from oracle.binding import Oracle
pkg = load_pkg(); ql = pkg.QwenLib(); orc = Oracle()
path_sel = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for shape, mode in (("tiny", "gauss"), ("tiny-untied", "gauss"), ("small", "gauss")):
    path = pkg.checkpoint.ensure_checkpoint("/tmp/qwen3_b200_ckpt", shape, seed=11, mode=mode)
    V = pkg.checkpoint.SHAPES[shape].vocab_size
    toks = np.random.default_rng(2).integers(0, V, size=5)
    with ql.open(path, 64) as gm, orc.open(path, 64, trace=True) as om:
        gm.set_path(path_sel)
        for pos, t in enumerate(toks):
            g, o = gm.forward(int(t), pos), om.forward(int(t), pos)
            d = np.abs(g - o)
            bad = d > 1e-2 + 1e-3 * np.abs(o)
            tr = om.trace()
            xs = gm.debug_read("x", gm.p.dim)
            hs = gm.debug_read("h", gm.p.hidden_dim)
            at = gm.debug_read("att", gm.p.n_heads * 128)
            print(f"{shape} pos {pos}: logit std {o.std():.3f} max|o| {np.abs(o).max():.3f} maxdiff {d.max():.3e} bad {bad.sum()} "
                  f"| att diff {np.abs(at - tr['att_out'][-1]).max():.2e} (max {np.abs(at).max():.2f}) "
                  f"h diff {np.abs(hs - tr['h'][-1]).max():.2e} (max {np.abs(hs).max():.2f}) nbad_h {(np.abs(hs - tr['h'][-1])>1e-4).sum()} nbad_att {(np.abs(at - tr['att_out'][-1])>1e-4).sum()} argmax {g.argmax()} {o.argmax()}", flush=True)
        L = gm.p.n_layers
        ok, ov = om.kv()
        for l in range(L):
            k, v = gm.kv_read(l, 0, len(toks))
            print(f"  layer {l} K diff {np.abs(k - ok[l,:len(toks)]).max():.2e} V diff {np.abs(v - ov[l,:len(toks)]).max():.2e} Kmax {np.abs(k).max():.2f} Vmax {np.abs(v).max():.2f}")
            if l > 3: break
