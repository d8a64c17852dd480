"""Oracle (CPU restatement) of the connected-region detection, pinned by the reference's own checks:
`validate_region_count` = count_regions == brute-force flood fill (split_detection.rs:490-560), the two unit
tests of extraction.rs:2586-2623, and structural properties of the labelling (split_detection.rs:662-893)."""
import numpy as np
import pytest

import helpers as H


def _object(oracle, graph, types=H.SAME0):
    gen = oracle.Generator(graph.nodes(), graph.root_node_id)
    return oracle.Object.generate(oracle.VoxelGenerator(gen, 1.0, types), 4)


def two_spheres_graph(gap=60.0, r1=25.0, r2=25.0):
    from impact_b200.graph import SDFGraph

    g = SDFGraph()  # should_split_off_disconnected_sphere (extraction.rs:2603-2623)
    a = g.sphere(r1)
    b = g.sphere(r2)
    b = g.translation(b, [gap, 0.0, 0.0])
    g.union(a, b, 1.0)
    return g


def check_labelling(obj, sd):
    """Every non-empty voxel carries a label below its chunk's region count, boundary regions are numbered
    first, face-adjacent non-empty voxels of one chunk share a label, and the roots partition the local
    regions exactly like the brute-force components."""
    ch, vx = obj.chunks(), obj.voxels()
    vox = vx.reshape(-1, 4096)
    lab = sd["voxel_labels"].reshape(-1, 4096)
    for c in np.nonzero(ch["kind"] == 2)[0]:
        o = ch["data_offset"][c]
        empty = (vox[o]["flags"] & 1) != 0
        l3 = lab[o].reshape(16, 16, 16)
        e3 = empty.reshape(16, 16, 16)
        assert np.all(lab[o][empty] == 255)
        rc, bc = sd["per_chunk"]["region_count"][c], sd["per_chunk"]["boundary_region_count"][c]
        if (~empty).any():
            assert lab[o][~empty].max() == rc - 1 and set(np.unique(lab[o][~empty])) == set(range(rc))
        else:
            assert rc == 0 and bc == 0
        boundary = np.ones((16, 16, 16), bool)
        boundary[1:-1, 1:-1, 1:-1] = False
        on_boundary = set(np.unique(l3[boundary & ~e3]))
        assert on_boundary == set(range(bc)), (c, on_boundary, bc)
        for ax in range(3):
            a = np.moveaxis(l3, ax, 0)
            ea = np.moveaxis(e3, ax, 0)
            both = ~ea[:-1] & ~ea[1:]
            assert np.array_equal(a[:-1][both], a[1:][both])


@pytest.mark.parametrize("graph,expected", [
    (lambda: H.box_graph(1.0), 1),             # connected_region_count_is_correct_for_single_voxel
    (lambda: H.sphere_graph(31.0), 1),
    (lambda: two_spheres_graph(), 2),           # should_split_off_disconnected_sphere
    (lambda: two_spheres_graph(60.0, 25.0, 9.0), 2),
    (lambda: H.complex_graph(0.5), 1),
])
def test_region_count_matches_brute_force(oracle, graph, expected):
    obj = _object(oracle, graph())
    sd = obj.split_detection()
    assert not sd["overflow"]
    assert sd["n_regions"] == obj.count_regions_brute_force() == expected
    assert sd["has_two"] == (expected >= 2)
    check_labelling(obj, sd)
    if expected >= 2:
        a, b = sd["two"]
        assert a != b and a < b  # roots are discovered in chunk order
        small = sd["candidates"][sd["smallest"]]
        other = sd["candidates"][1 - sd["smallest"]]
        assert (small["non_uniform_chunk_count"], small["chunk_count"]) <= (other["non_uniform_chunk_count"], other["chunk_count"])


def test_smaller_sphere_is_the_one_extracted(oracle):
    obj = _object(oracle, two_spheres_graph(60.0, 25.0, 9.0))
    sd = obj.split_detection()
    small = sd["candidates"][sd["smallest"]]
    # the r = 9 sphere sits at +x: its chunk range lies beyond the big sphere's
    assert small["chunk_count"] < sd["candidates"][1 - sd["smallest"]]["chunk_count"]
    assert small["chunk_min"][0] >= sd["candidates"][1 - sd["smallest"]]["chunk_min"][0]


def test_noisy_object_with_debris_matches_brute_force(oracle):
    # strong noise on a small sphere sheds detached blobs: many regions, interior-only regions included
    g = H.noisy_sphere_graph(20.0, 4)
    obj = _object(oracle, g)
    sd = obj.split_detection()
    assert not sd["overflow"]
    assert sd["n_regions"] == obj.count_regions_brute_force()
    check_labelling(obj, sd)


def test_absorption_can_split_an_object(oracle):
    # a dumbbell: two spheres joined by a thin capsule; absorbing the bridge disconnects them
    from impact_b200.graph import SDFGraph

    g = SDFGraph()
    a = g.sphere(14.0)
    b = g.translation(g.sphere(14.0), [44.0, 0.0, 0.0])
    bridge = g.capsule(30.0, 3.0)
    bridge = g.rotation_from_axis_angle(bridge, [0.0, 0.0, 1.0], float(np.pi / 2))
    bridge = g.translation(bridge, [22.0, 0.0, 0.0])
    g.union(g.union(a, b, 1.0), bridge, 1.0)
    obj = _object(oracle, g)
    assert obj.split_detection()["n_regions"] == obj.count_regions_brute_force() == 1
    shape = np.array(obj.info()["chunk_counts"]) * 16
    center = np.float32([0.5 * shape[0], 0.5 * shape[1], 0.5 * shape[2]])
    for _ in range(3):
        obj.absorb_sphere(center, 7.0, 9.0)
    sd = obj.split_detection()
    assert sd["n_regions"] == obj.count_regions_brute_force() == 2
    check_labelling(obj, sd)
