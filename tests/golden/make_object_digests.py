"""Writes tests/golden/object_digests.json: sha256 digests of the voxel object and mesh the CPU oracle produces for a
few graphs of tests/helpers.py (GOLDEN_OBJECTS). The oracle is checked against them on every CPU test run (so a change
of the restatement is noticed) and the CUDA path against the same committed digests on the GPU.

    python tests/golden/make_object_digests.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers as H  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402

out = {}
for name, (make, types_name) in H.GOLDEN_OBJECTS.items():
    g = make()
    types = getattr(H, types_name)
    gen = O.Generator(g.nodes(), g.root_node_id)
    obj = O.Object.generate(O.VoxelGenerator(gen, 1.0, types), 4)
    m = obj.mesh(2)
    out[name] = {"types": types_name, "chunk_counts": [int(x) for x in obj.info()["chunk_counts"]],
                 "object": H.object_digest(obj.chunks(), obj.voxels()),
                 "mesh": H.mesh_digest(m.positions, m.normals, m.indices, m.index_materials, m.submeshes, m.vertex_ranges),
                 "vertices": int(m.n_vertices), "indices": int(m.n_indices)}
with open(os.path.join(HERE, "object_digests.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
