"""CPU: the general mesh → multi-scale graph ingest (utils/mesh_ingest.py) on the structured tri(nx, ny) geometry must
reproduce utils.synthetic.make_tri_mesh's integer artefacts bit for bit; containment on a perturbed (non-nested) mesh
against a brute-force point-in-polygon; the neutral .npz format round-trips; the Morton renumbering is a permutation that
keeps the graph isomorphic."""
import numpy as np
import pytest
import torch

import mswe_gnn_b200  # noqa: F401
from mswe_gnn_b200.utils import mesh_ingest as MI
from mswe_gnn_b200.utils.synthetic import make_tri_mesh


@pytest.mark.parametrize("nx,ny,S", [(8, 8, 3), (16, 8, 4), (4, 12, 2), (6, 5, 1)])
def test_structured_mesh_through_general_ingest_is_bit_exact(nx, ny, S):
    ref = make_tri_mesh(nx, ny, S, previous_t=3)
    got = MI.build_multiscale_graph(MI.structured_tri_levels(nx, ny, S), inflow_xy=(0.5, 0.0), previous_t=3)
    for k in ("edge_index", "node_ptr", "edge_ptr", "intra_mesh_edge_index", "intra_edge_ptr", "node_BC"):
        assert torch.equal(getattr(got, k), getattr(ref, k)), k
    assert got.x.shape == ref.x.shape and got.BC.shape == ref.BC.shape and got.edge_attr.shape == ref.edge_attr.shape
    # geometry-derived attributes: triangle areas and centre distances of the unit grid
    n0 = int(ref.node_ptr[1]) - 1
    assert torch.allclose(got.x[:n0, 0], torch.full((n0,), 0.5))


def test_containment_matches_brute_force_on_a_non_nested_mesh():
    rng = np.random.default_rng(0)
    lv = MI.structured_tri_levels(12, 10, 2)
    for l in lv:                                                  # jitter the interior nodes: containment is no tree any more
        xy = l["node_xy"].copy()
        inner = (xy[:, 0] > 0) & (xy[:, 0] < 12) & (xy[:, 1] > 0) & (xy[:, 1] < 10)
        xy[inner] += rng.uniform(-0.3, 0.3, size=(int(inner.sum()), 2))
        l["node_xy"] = xy
    ctr = MI.face_centres(lv[0]["node_xy"], lv[0]["face_nodes"])
    got = MI.containment_edges(lv[1]["node_xy"], lv[1]["face_nodes"], ctr)
    P = lv[1]["node_xy"][lv[1]["face_nodes"]]

    def inside(tri, p):
        d = [(tri[(k + 1) % 3][0] - tri[k][0]) * (p[1] - tri[k][1]) - (tri[(k + 1) % 3][1] - tri[k][1]) * (p[0] - tri[k][0]) for k in range(3)]
        return all(v > 0 for v in d) or all(v < 0 for v in d)
    ref = [(c, f) for c in range(P.shape[0]) for f in range(ctr.shape[0]) if inside(P[c], ctr[f])]
    assert got.T.tolist() == [list(p) for p in ref]
    deg = np.bincount(got[1], minlength=ctr.shape[0])
    assert deg.max() == 1 and got.shape[1] == ctr.shape[0]       # triangles tile the plane: every centre in exactly one


def test_npz_round_trip_and_morton_renumbering(tmp_path):
    lv = MI.structured_tri_levels(8, 8, 3)
    MI.save_levels(str(tmp_path / "mesh.npz"), lv)
    back = MI.load_levels(str(tmp_path / "mesh.npz"))
    for a, b in zip(lv, back):
        assert np.array_equal(a["node_xy"], b["node_xy"]) and np.array_equal(a["face_nodes"], b["face_nodes"])
    g0 = MI.build_multiscale_graph(lv, (0.5, 0.0), previous_t=3)
    g1 = MI.build_multiscale_graph(lv, (0.5, 0.0), previous_t=3, renumber=True)
    assert torch.equal(g0.node_ptr, g1.node_ptr) and torch.equal(g0.edge_ptr, g1.edge_ptr) and torch.equal(g0.intra_edge_ptr, g1.intra_edge_ptr)
    # same multiset of edge lengths / areas, same degree sequence: the renumbered graph is the same mesh
    assert torch.equal(torch.sort(g0.edge_attr[:, 0]).values, torch.sort(g1.edge_attr[:, 0]).values)
    d0 = torch.bincount(g0.edge_index[1], minlength=g0.x.shape[0])
    d1 = torch.bincount(g1.edge_index[1], minlength=g1.x.shape[0])
    assert torch.equal(torch.sort(d0).values, torch.sort(d1).values)
    # and it is local: consecutive faces of the finest level are close in space
    n0 = int(g1.node_ptr[1]) - 1
    step1 = (g1.pos[1:n0] - g1.pos[:n0 - 1]).norm(dim=1).mean()
    assert float(step1) < 2.0
    perm = MI.morton_order(np.random.default_rng(1).uniform(size=(1000, 2)))
    assert sorted(perm.tolist()) == list(range(1000))
