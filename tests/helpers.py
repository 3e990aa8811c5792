"""Shared test plumbing: fixture loading, input regeneration, model construction."""
import hashlib
import json
import os

import numpy as np
import torch

import mswe_gnn_b200  # noqa: F401  (import shim)
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh
from oracle import swe_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_CONFIG_MODELS = dict(hid_features=64, mlp_layers=3, seed=666, learned_residuals=True, mlp_activation="prelu",
                         gnn_activation="tanh", edge_mlp=True, normalize=True, with_filter_matrix=True,
                         with_gradient=True, with_WL=True, K=4, learned_pooling=False, skip_connections=True)
# == /root/reference/config.yaml:42-58 minus model_type (tests/test_oracle_pinned.py checks the copy)


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return meta, z


def sd_sha(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def make_mesh(meta):
    mesh = dict(meta["mesh"])
    if meta["model"] == "GNN":
        return make_single_scale_mesh(**mesh)
    return make_tri_mesh(**mesh)


def build_model(meta, z=None, device="cpu"):
    """Our model built from the fixture's constructor arguments (+ stored weights if any)."""
    from mswe_gnn_b200.models.gnn import GNN, MSGNN
    cls = MSGNN if meta["model"] == "MSGNN" else GNN
    m = cls(**meta["ctor"])
    if z is not None:
        w = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
        if w:
            m.load_state_dict(w)
    return m.to(device)


def spec_of(meta):
    c = dict(meta["ctor"])
    return O.ModelSpec(meta["model"], **c)


def to_device(data, device):
    return data.to(device)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def assert_close_masked(ours, ref, rtol, atol, what="", flip_band=2e-5, max_flip_frac=2e-3):
    """|ours - ref| <= atol + rtol*|ref| except at wet/dry flips: outputs pass through
    relu and the |h| > 1e-4 dry mask (models.py:79-91), which are discontinuous, so an element whose
    reference depth sits within `flip_band` of a threshold may legitimately land on the other
    side; such rows are counted and bounded instead."""
    ours, ref = ours.double().cpu(), ref.double().cpu()
    err = (ours - ref).abs()
    ok = err <= atol + rtol * ref.abs()
    if ok.all():
        return 0
    bad_rows = (~ok).reshape(ok.shape[0], -1).any(1)
    # candidates for a flip: one side exactly zero, the other tiny or near the 1e-4 threshold
    h_ref = ref.reshape(ref.shape[0], -1)[:, 0:1] if ref.dim() == 2 else ref[:, 0]
    h_our = ours.reshape(ours.shape[0], -1)[:, 0:1] if ours.dim() == 2 else ours[:, 0]
    near = ((h_ref.abs() - 1e-4).abs() < flip_band) | ((h_our.abs() - 1e-4).abs() < flip_band) | \
           (h_ref.abs() < flip_band) | (h_our.abs() < flip_band)
    near = near.reshape(near.shape[0], -1).any(1)
    unexplained = bad_rows & ~near
    assert not unexplained.any(), \
        f"{what}: {int(unexplained.sum())} rows differ beyond tolerance (max err {float(err.max()):.3e})"
    frac = float(bad_rows.float().mean())
    assert frac <= max_flip_frac, f"{what}: {frac:.4%} rows flipped wet/dry (> {max_flip_frac:.4%})"
    return int(bad_rows.sum())
