"""Checkpoint interop: the reference saves `tf.train.Saver` checkpoints whose tensors are the graph
variables created by agents/policies.py / agents/utils.py:

    <policy>_<i>a/{pi,v}_fcw/{w,b}      fc over the wave block            agents/policies.py:107-110, utils.py:66-74
    <policy>_<i>a/{pi,v}_fcf/{w,b}      fc over the fingerprints (MA2C)   agents/policies.py:202-203
    <policy>_<i>a/{pi,v}_fct/{w,b}      fc over the wait block (n_w > 0)  agents/policies.py:110,207
    <policy>_<i>a/{pi,v}_lstm/{wx,wh,b} LSTM, gates i,f,o,u               agents/utils.py:95-100
    <policy>_<i>a/{pi,v}_fc/{w,b}       FcACPolicy hidden layer           agents/policies.py:236
    <policy>_<i>a/{pi,v}/{w,b}          softmax / value head              agents/policies.py:18-26
    <policy> = 'lstm' (IA2C), 'fplstm' (MA2C), 'fc' (FcACPolicy); rows of `wx` follow concat(fcw, fcf, fct).
    IQL:  lr_<i>a_q/q/{w,b};  dqn_<i>a_q/{q_fcw,q_fct,q_fc_0,q}/{w,b}      agents/policies.py:297-301,341-389

The Saver is created before the optimizer (agents/models.py:154-159), so a reference checkpoint holds exactly these
weights and no RMSProp / Adam slots.  `export_named` / `import_named` map between that name space and the flat
parameter vector of `PolicyLayout`; the name list itself is pinned by tests/golden/learner_*.npz, which was written by
the reference's own graph builders.  On disk we use `checkpoint-<step>.npz` keyed by those names (TensorFlow is not
available here to write its bundle format; scripts/convert_tf_checkpoint.py converts either way on a TF host).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .layout import PolicyLayout

NETS = ("pi", "v")


def policy_prefix(layout: PolicyLayout) -> str:
    if not layout.recurrent:
        return "fc"
    return "fplstm" if layout.ff > 0 else "lstm"


def variable_specs(layout: PolicyLayout, prefix: str | None = None) -> List[Tuple[str, str, int, tuple]]:
    """[(tf name, view key, unit, shape)] in the reference's graph-construction order."""
    prefix = prefix or policy_prefix(layout)
    L = layout
    out = []
    for a in range(L.A):
        for k, net in enumerate(NETS):
            u = 2 * a + k
            base = "%s_%da/%s" % (prefix, a, net)
            nw, nf, nt = int(L.n_wave[a]), int(L.n_fp[a]), int(L.n_wait[a])
            out.append((base + "_fcw/w", "fcw_w%d" % u, u, (nw, L.fw)))
            out.append((base + "_fcw/b", "fcw_b%d" % u, u, (L.fw,)))
            if L.ff > 0:
                out.append((base + "_fcf/w", "fcf_w%d" % u, u, (nf, L.ff)))
                out.append((base + "_fcf/b", "fcf_b%d" % u, u, (L.ff,)))
            if L.ft > 0 and nt > 0:
                out.append((base + "_fct/w", "fct_w%d" % u, u, (nt, L.ft)))
                out.append((base + "_fct/b", "fct_b%d" % u, u, (L.ft,)))
            if L.recurrent:
                out.append((base + "_lstm/wx", "wx", u, (L.dx, 4 * L.h)))
                out.append((base + "_lstm/wh", "wh", u, (L.h, 4 * L.h)))
                out.append((base + "_lstm/b", "bl", u, (4 * L.h,)))
            else:
                out.append((base + "_fc/w", "wx", u, (L.dx, L.h)))
                out.append((base + "_fc/b", "bl", u, (L.h,)))
            n_out = int(L.n_a[a]) if k == 0 else 1
            out.append((base + "/w", "wo", u, (L.h, n_out)))
            out.append((base + "/b", "bo", u, (n_out,)))
    return out


def export_named(layout: PolicyLayout, flat: np.ndarray, prefix: str | None = None) -> Dict[str, np.ndarray]:
    v = layout.views(np.asarray(flat))
    named = {}
    for name, key, u, shape in variable_specs(layout, prefix):
        if key in ("wx", "wh", "bl"):
            arr = v[key][u]
        elif key == "wo":
            arr = v["wo"][u][:, :shape[1]]
        elif key == "bo":
            arr = v["bo"][u][:shape[0]]
        else:
            arr = v[key]
        named[name] = np.array(arr, dtype=np.float32).reshape(shape)
    return named


def import_named(layout: PolicyLayout, named: Dict[str, np.ndarray], prefix: str | None = None,
                 dtype=np.float32) -> np.ndarray:
    """Flat parameter vector from {tf name: array}; raises KeyError / ValueError on a missing or mis-shaped tensor.
    Padded head columns (n_a < max_na, value heads) are zero."""
    flat = np.zeros(layout.n_params, dtype)
    v = layout.views(flat)
    for name, key, u, shape in variable_specs(layout, prefix):
        if name not in named:
            raise KeyError("checkpoint has no tensor %r" % name)
        arr = np.asarray(named[name])
        if tuple(arr.shape) != tuple(shape):
            raise ValueError("tensor %r has shape %s, expected %s" % (name, arr.shape, shape))
        if key in ("wx", "wh", "bl"):
            v[key][u][...] = arr
        elif key == "wo":
            v["wo"][u][:, :shape[1]] = arr
        elif key == "bo":
            v["bo"][u][:shape[0]] = arr
        else:
            v[key][...] = arr
    return flat


def save_npz(path: str, named: Dict[str, np.ndarray], extra: Dict[str, np.ndarray] | None = None) -> None:
    """`extra` entries (optimizer slots, step) are stored under '__b200__/' and ignored by a TF-side converter."""
    out = dict(named)
    for k, a in (extra or {}).items():
        out["__b200__/" + k] = np.asarray(a)
    np.savez(path, **out)


def load_npz(path: str) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    z = np.load(path)
    named = {k: z[k] for k in z.files if not k.startswith("__b200__/")}
    extra = {k[len("__b200__/"):]: z[k] for k in z.files if k.startswith("__b200__/")}
    return named, extra
