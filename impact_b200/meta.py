"""Meta SDF graph compile: file readers, the binding of the library's `ivx_meta_compile` (csrc/meta.cpp, the product path:
`compile_meta_nodes`, `compile_file`, `compile_graph_file`) and `MetaCompiler`, a Python mirror of the same compile that
the tests hold the C++ against node for node (tests/test_meta_native.py).

`MetaSDFGraph::build_in(scale_factor, seed) -> SDFGraph`
(engine/crates/impact_voxel/src/generation/sdf/meta.rs:741-896, node resolvers :1194-2260, parameter
sampling meta/params.rs:85-264, stable seeding meta.rs:993-1097 + impact_math/src/random/splitmix.rs:4-20).

The compile stays on the host, like in the reference; its only compute-heavy part — the signed-distance
probes of the surface-seeking nodes (sphere cast, meta.rs:2534-2748) — runs on the GPU through
`ivx_program_eval_blocks`, batched over all instances of a node (the reference probes one instance at a
time on the CPU). All arithmetic is numpy float32 in the reference's operation order.

PARITY NOTE: the reference draws its random numbers from the third-party crate `fastrand` 2.3.0
(engine/Cargo.lock:922-923, not under /root/reference). `Rng` below restates its published wyrand
generator and Lemire range reduction; no reference test pins RNG output, so the compiled
asteroid is "an asteroid from the reference's graph", not provably the same instance the Rust build
produces. Everything downstream (generation, meshing) is checked bit-exactly against the oracle on the
atomic graph this module emits.

Implemented node kinds — all 21 of meta.rs: Points, Spheres, Capsules, Boxes, Translation, Rotation, Scaling, Similarity,
StratifiedGridTransforms, SphereSurfaceTransforms, ClosestTranslationToSurface, RayTranslationToSurface,
RotationToGradient, StochasticSelection, SDFInstantiation, TransformApplication, MultifractalNoiseSDFModifier,
SDFUnion, SDFSubtraction, SDFIntersection, SDFGroupUnion. The three that probe an SDF (ClosestTranslationToSurface,
RayTranslationToSurface, RotationToGradient) evaluate it on the device, all instances of the node in one batch.
"""
from __future__ import annotations

import math
import os
import re

import numpy as np

from .graph import SDFGraph

f32 = np.float32
M64 = (1 << 64) - 1

# ------------------------------------------------------------------------------------------------
# RON subset reader (engine/crates/impact_voxel/src/generation/import.rs:51-72 uses the `ron` crate)


class Tagged:
    """`Name(...)`: an enum variant / named struct. fields is a dict (named), a list (positional) or None."""

    def __init__(self, tag, fields):
        self.tag, self.fields = tag, fields

    def __getitem__(self, k):
        return self.fields[k]

    def __repr__(self):
        return f"{self.tag}({self.fields})"


_TOKEN = re.compile(r"\s*(?:(//[^\n]*)|([A-Za-z_][A-Za-z_0-9]*)|(-?\d+\.\d*(?:[eE][-+]?\d+)?|-?\d+(?:[eE][-+]?\d+)?)|(\"(?:[^\"\\]|\\.)*\")|(.))")


def _tokenize(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            break
        pos = m.end()
        if m.group(1):
            continue
        if m.group(2):
            out.append(("id", m.group(2)))
        elif m.group(3):
            out.append(("num", m.group(3)))
        elif m.group(4):
            out.append(("str", m.group(4)[1:-1]))
        elif m.group(5) and not m.group(5).isspace():
            out.append(("p", m.group(5)))
    return out


def parse_ron(text):
    toks = _tokenize(text)
    i = 0

    def peek():
        return toks[i] if i < len(toks) else ("eof", "")

    def eat(kind=None, val=None):
        nonlocal i
        t = peek()
        if (kind and t[0] != kind) or (val and t[1] != val):
            raise ValueError(f"RON: expected {kind} {val}, got {t} at token {i}")
        i += 1
        return t

    def group(close):
        """contents of (...) : named fields → dict, positional → list"""
        named, items = None, []
        while peek() != ("p", close):
            if peek()[0] == "id" and i + 1 < len(toks) and toks[i + 1] == ("p", ":"):
                k = eat("id")[1]
                eat("p", ":")
                v = value()
                if named is None:
                    named = {}
                named[k] = v
            else:
                items.append(value())
            if peek() == ("p", ","):
                eat()
        eat("p", close)
        return named if named is not None else items

    def value():
        t = peek()
        if t[0] == "num":
            eat()
            return float(t[1]) if any(c in t[1] for c in ".eE") else int(t[1])
        if t[0] == "str":
            eat()
            return t[1]
        if t == ("p", "["):
            eat()
            return group("]")
        if t == ("p", "("):
            eat()
            g = group(")")
            return g[0] if isinstance(g, list) and len(g) == 1 else g
        if t[0] == "id":
            eat()
            if t[1] in ("true", "false"):
                return t[1] == "true"
            if peek() == ("p", "("):
                eat()
                g = group(")")
                if isinstance(g, list) and len(g) == 1:
                    g = g[0]  # newtype variant
                return Tagged(t[1], g)
            return Tagged(t[1], None)
        raise ValueError(f"RON: unexpected token {t}")

    return value()


def load_vgen_ron(path):
    """`*.vgen.ron` → list of meta nodes (`VoxelGenerator { sdf_graph: MetaSDFGraph { nodes } }`)."""
    with open(path) as f:
        doc = parse_ron(f.read())
    return doc["sdf_graph"]["nodes"]


# ------------------------------------------------------------------------------------------------
# the editor's `*.graph.ron` (apps/voxel_generator/src/editor/meta/io.rs:15-70): nodes with positional parameters and
# links, turned into the same meta nodes by the editor's own rules (build.rs:53-100, node_kind.rs `build` of every kind)

_SAMPLING = ("OnlyOnce", "PerInstance")            # "Only once", "Per instance" / "Per SDF" (meta.rs:2272-2280)
_COMPOSITION = ("Post", "Pre")                      # node_kind.rs:446
_ROTATION = ("Identity", "RadialOutwards", "RadialInwards")  # node_kind.rs:956
_ANCHOR = ("Origin", "ShapeBoundaryAtOrigin")       # node_kind.rs:1036
# kind → (child field names in slot order, [(field, how)]): how = "dist" (distributed → ParamSpec), "num" (UInt / Float
# value) or a tuple of enum variant names indexed by the stored variant number
EDITOR_NODE_KINDS = {
    "Points": ((), [("count", "num")]),
    "Spheres": ((), [("radius", "dist"), ("center_x", "dist"), ("center_y", "dist"), ("center_z", "dist"), ("count", "num"),
                     ("seed", "num"), ("sampling", _SAMPLING)]),
    "Capsules": ((), [("segment_length", "dist"), ("radius", "dist"), ("center_x", "dist"), ("center_y", "dist"),
                      ("center_z", "dist"), ("count", "num"), ("seed", "num"), ("sampling", _SAMPLING)]),
    "Boxes": ((), [("extent_x", "dist"), ("extent_y", "dist"), ("extent_z", "dist"), ("center_x", "dist"), ("center_y", "dist"),
                   ("center_z", "dist"), ("count", "num"), ("seed", "num"), ("sampling", _SAMPLING)]),
    "Translation": (("child_id",), [("composition", _COMPOSITION), ("translation_x", "dist"), ("translation_y", "dist"),
                                    ("translation_z", "dist"), ("seed", "num"), ("sampling", _SAMPLING)]),
    "Rotation": (("child_id",), [("composition", _COMPOSITION), ("tilt_angle", "dist"), ("turn_angle", "dist"),
                                 ("roll_angle", "dist"), ("seed", "num"), ("sampling", _SAMPLING)]),
    "Scaling": (("child_id",), [("composition", _COMPOSITION), ("scaling", "dist"), ("seed", "num"), ("sampling", _SAMPLING)]),
    "Similarity": (("child_id",), [("composition", _COMPOSITION), ("scale", "dist"), ("tilt_angle", "dist"), ("turn_angle", "dist"),
                                   ("roll_angle", "dist"), ("translation_x", "dist"), ("translation_y", "dist"),
                                   ("translation_z", "dist"), ("seed", "num"), ("sampling", _SAMPLING)]),
    "StratifiedGridTransforms": (("child_id",), [("shape_x", "dist"), ("shape_y", "dist"), ("shape_z", "dist"),
                                                 ("cell_extent_x", "dist"), ("cell_extent_y", "dist"), ("cell_extent_z", "dist"),
                                                 ("jitter_fraction", "dist"), ("seed", "num")]),
    "SphereSurfaceTransforms": (("child_id",), [("radius", "dist"), ("jitter_fraction", "dist"), ("rotation", _ROTATION),
                                                ("seed", "num")]),
    "ClosestTranslationToSurface": (("surface_sdf_id", "subject_id"), []),
    "RayTranslationToSurface": (("surface_sdf_id", "subject_id"), [("anchor", _ANCHOR)]),
    "RotationToGradient": (("gradient_sdf_id", "subject_id"), []),
    "StochasticSelection": (("child_id",), [("min_pick_count", "num"), ("max_pick_count", "num"), ("pick_probability", "num"),
                                            ("seed", "num")]),
    "SDFInstantiation": (("child_id",), []),
    "TransformApplication": (("sdf_id", "instance_id"), []),
    "MultifractalNoiseSDFModifier": (("child_id",), [("octaves", "dist"), ("frequency", "dist"), ("lacunarity", "dist"),
                                                     ("persistence", "dist"), ("amplitude", "dist"), ("seed", "num"),
                                                     ("sampling", _SAMPLING)]),
    "SDFUnion": (("child_1_id", "child_2_id"), [("smoothness", "num")]),
    "SDFSubtraction": (("child_1_id", "child_2_id"), [("smoothness", "num")]),
    "SDFIntersection": (("child_1_id", "child_2_id"), [("smoothness", "num")]),
    "SDFGroupUnion": (("child_id",), [("smoothness", "num")]),
}


def _payload(t):
    """`Tag(x)` → x (the parser keeps a single positional field as the value or as a one-element list)."""
    return t.fields[0] if isinstance(t.fields, list) else t.fields


_EDITOR_DISCRETE = {"shape_x", "shape_y", "shape_z", "octaves"}  # DiscreteParamSpec fields (meta.rs)


def _editor_source(src, discrete=False):
    """`ValueSource` → `ContValueSource` / `DiscreteValueSource` (param.rs:883-905; `fixed as u32` for discrete ones)."""
    if src["variant"].tag == "Fixed":
        return Tagged("Fixed", int(max(src["fixed"], 0.0)) if discrete else src["fixed"])
    fp = src["from_param"]
    lin = fp["mapping"]["linear"]
    return Tagged("FromParam", {"idx": int(fp["param_idx"]),
                                "mapping": Tagged("Linear", {"offset": lin["offset"], "scale": lin["scale"]})})


def _editor_spec(dist, discrete=False):
    """`ParamDistribution` → `ContParamSpec` / `DiscreteParamSpec` (param.rs:751-785)."""
    v = dist["variant"].tag
    some = _payload  # Some((…))
    if v == "Constant":
        return Tagged("Constant", _editor_source(dist["constant"], discrete))
    if v == "Uniform":
        u = some(dist["uniform"])
        return Tagged("Uniform", {"min": _editor_source(u["min"], discrete), "max": _editor_source(u["max"], discrete)})
    if discrete:
        raise ValueError(f"distribution {v} is not available for discrete parameters")
    if v == "UniformCosAngle":
        u = some(dist["uniform_cos_angle"])
        return Tagged("UniformCosAngle", {"min_angle": _editor_source(u["min_angle"]), "max_angle": _editor_source(u["max_angle"])})
    if v == "PowerLaw":
        u = some(dist["power_law"])
        return Tagged("PowerLaw", {"min": _editor_source(u["min"]), "max": _editor_source(u["max"]),
                                   "exponent": _editor_source(u["exponent"])})
    raise ValueError(f"unknown distribution variant {v}")


def load_graph_ron(path):
    """The voxel generator editor's `*.graph.ron` → (meta nodes, voxel_extent, scale_factor, seed).

    Restates `build_meta_graph` (apps/voxel_generator/src/editor/meta/build.rs:53-100): the graph below the Output node
    (id 0; its parameters are voxel extent, scale factor and seed, node_kind.rs `get_properties_from_output_node`) is
    walked depth first, first child first, and every node is added after its children, so meta node ids are post-order
    positions and the root is the last node. Raises ValueError for a graph without output or with an unattached slot
    (`build_meta_graph` returns None there)."""
    with open(path) as f:
        return graph_ron_nodes(f.read())


def graph_ron_nodes(text):
    """`load_graph_ron` on the file's text."""
    doc = parse_ron(text)
    by_id = {int(n["id"]): n for n in doc["nodes"]}

    def link(l):
        if l.tag != "Some":
            raise ValueError("unattached child slot")
        return int(_payload(l)["to_node"])

    if doc["kind"].tag == "Subgraph":
        # a `*.subgraph.ron` fragment (io.rs:31-39): no Output node; its root is named by the file and the output
        # properties take the editor's defaults for a new Output node (node_kind.rs:101-104, 1727-1770)
        root = int(doc["kind"]["root_node_id"])
        if root not in by_id:
            raise ValueError("subgraph root node is missing")
        voxel_extent, scale_factor, seed = 0.25, 0.25, 0
    else:
        out = by_id.get(0)
        if out is None or out["kind"].tag != "Output":
            raise ValueError("graph has no Output node")
        p = out["params"]
        voxel_extent, scale_factor, seed = float(_payload(p[0])), float(_payload(p[1])), int(_payload(p[2]))
        root = link(out["links_to_children"][0])
    id_map, nodes = {}, []
    stack = [("visit", root)]
    while stack:
        op, nid = stack.pop()
        node = by_id[nid]
        if op == "visit":
            if nid in id_map:
                continue
            stack.append(("build", nid))
            for l in reversed(node["links_to_children"]):
                stack.append(("visit", link(l)))
            continue
        if nid in id_map:
            continue
        kind = node["kind"].tag
        child_names, params = EDITOR_NODE_KINDS[kind]
        if len(node["params"]) != len(params):
            raise ValueError(f"{kind} node {nid}: expected {len(params)} parameters, found {len(node['params'])}")
        fields = {}
        for name, l in zip(child_names, node["links_to_children"]):
            fields[name] = id_map[link(l)]
        for (name, how), val in zip(params, node["params"]):
            if how == "dist":
                fields[name] = _editor_spec(_payload(val), name in _EDITOR_DISCRETE)
            elif how == "num":
                fields[name] = _payload(val)
            else:
                fields[name] = Tagged(how[int(_payload(val))], None)
        id_map[nid] = len(nodes)
        nodes.append(Tagged(kind, fields))
    return nodes, voxel_extent, scale_factor, seed


def compile_graph_file(path, ctx=None):
    """`build_sdf_graph` (build.rs:102-128): the editor file compiled with its own scale factor and seed →
    (atomic SDFGraph, voxel_extent)."""
    nodes, voxel_extent, scale_factor, seed = load_graph_ron(path)
    return compile_meta_nodes(nodes, scale_factor, seed, ctx), voxel_extent


# ------------------------------------------------------------------------------------------------
# randomness: splitmix (impact_math/src/random/splitmix.rs) + fastrand wyrand (restated, see header)

def splitmix(state):
    state = (state + 0x9E3779B97F4A7C15) & M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def splitmix2(a, b):
    return splitmix(a ^ splitmix(b))


def splitmix3(a, b, c):
    return splitmix2(splitmix2(a, b), c)


class Rng:
    def __init__(self, seed):
        self.s = seed & M64

    def gen_u64(self):
        self.s = (self.s + 0x2D358DCCAA6C78A5) & M64
        t = self.s * (self.s ^ 0x8BB84B93962EACC9)
        return (t & M64) ^ (t >> 64)

    def gen_u32(self):
        return self.gen_u64() & 0xFFFFFFFF

    def _mod(self, n, bits):
        mask = (1 << bits) - 1
        gen = self.gen_u32 if bits == 32 else self.gen_u64
        r = gen()
        hi, lo = (r * n) >> bits, (r * n) & mask
        if lo < n:
            t = ((-n) & mask) % n
            while lo < t:
                r = gen()
                hi, lo = (r * n) >> bits, (r * n) & mask
        return hi

    def u32_inclusive(self, lo, hi):
        if lo == 0 and hi == 0xFFFFFFFF:
            return self.gen_u32()
        return (lo + self._mod((hi - lo + 1) & 0xFFFFFFFF, 32)) & 0xFFFFFFFF

    def usize_inclusive(self, lo, hi):
        return (lo + self._mod((hi - lo + 1) & M64, 64)) & M64

    def f32(self):
        bits = np.uint32(0x3F800000 + (self.gen_u32() >> 9))
        return bits.view(f32) - f32(1.0)

    def f32_in_range(self, start, end):
        t = self.f32()
        return f32(start) + t * (f32(end) - f32(start))

    def clone_random_subset(self, count, source):
        """`Rng::clone_random_subset_from_slice` (impact_math/src/random.rs:58-93): reservoir sampling."""
        dest = list(source[: min(count, len(source))])
        if count == 0 or count >= len(source):
            return dest
        idx = count
        for item in source[count:]:
            x = self.usize_inclusive(0, idx)
            if x < count:
                dest[x] = item
            idx += 1
        return dest


# ------------------------------------------------------------------------------------------------
# f32 vector / quaternion helpers (glam SSE2 operation order where it is known, see oracle_math.hpp)

def v3(x, y, z):
    return np.array([x, y, z], f32)


def dot(a, b):
    return f32(f32(a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])


def norm(a):
    return f32(np.sqrt(dot(a, a)))


def cross(l, r):
    return v3(l[1] * r[2] - l[2] * r[1], l[2] * r[0] - l[0] * r[2], l[0] * r[1] - l[1] * r[0])


QID = np.array([0, 0, 0, 1], f32)


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], f32)


def quat_rotate(q, v):
    b = q[:3]
    w = q[3]
    b2 = dot(b, b)
    t1 = v * f32(w * w - b2)
    t2 = b * f32(dot(v, b) * f32(2.0))
    t3 = cross(b, v) * f32(w * f32(2.0))
    return (t1 + t2 + t3).astype(f32)


def quat_conj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]], f32)


def quat_from_axis_angle(axis, angle):
    half = f32(angle) * f32(0.5)
    s, c = f32(math.sin(float(half))), f32(math.cos(float(half)))
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, c], f32)


def any_orthonormal_vector(v):
    # glam Vec3::any_orthonormal_vector
    sign = f32(math.copysign(1.0, float(v[2])))
    a = f32(-1.0) / (sign + v[2])
    b = v[0] * v[1] * a
    return v3(b, sign + v[1] * v[1] * a, -v[1])


def quat_from_rotation_arc(frm, to):
    # glam Quat::from_rotation_arc (UnitQuaternion::rotation_between_axes, quaternion.rs:345-350)
    one_minus_eps = f32(1.0) - f32(2.0) * np.finfo(f32).eps
    d = dot(frm, to)
    if d > one_minus_eps:
        return QID.copy()
    if d < -one_minus_eps:
        return quat_from_axis_angle(any_orthonormal_vector(frm), f32(math.pi))
    c = cross(frm, to)
    q = np.array([c[0], c[1], c[2], f32(1.0) + d], f32)
    n = f32(np.sqrt(f32(f32(f32(q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3])))
    return (q / n).astype(f32)


class Sim:
    """`Similarity3` (impact_math/src/transform/similarity.rs:22-26): scaling, then rotation, then translation."""

    __slots__ = ("t", "r", "s")

    def __init__(self, t=None, r=None, s=1.0):
        self.t = v3(0, 0, 0) if t is None else np.asarray(t, f32)
        self.r = QID.copy() if r is None else np.asarray(r, f32)
        self.s = f32(s)

    def translated(self, t):
        return Sim(self.t + t, self.r, self.s)

    def rotated(self, q):
        return Sim(quat_rotate(q, self.t), quat_mul(q, self.r), self.s)

    def scaled(self, s):
        return Sim(f32(s) * self.t, self.r, f32(s) * self.s)

    def applied_to_translation(self, t):
        return Sim(quat_rotate(self.r, self.s * t) + self.t, self.r, self.s)

    def applied_to_rotation(self, q):
        return Sim(self.t, quat_mul(self.r, q), self.s)

    def applied_to_scaling(self, s):
        return Sim(self.t, self.r, self.s * f32(s))

    def mul(self, b):  # self * b
        return Sim(quat_rotate(self.r, self.s * b.t) + self.t, quat_mul(self.r, b.r), self.s * b.s)

    def transform_point(self, p):
        return (quat_rotate(self.r, self.s * p) + self.t).astype(f32)

    def transform_vector(self, v):
        return quat_rotate(self.r, self.s * v)

    def inverse_transform_point(self, p):
        return (quat_rotate(quat_conj(self.r), (p - self.t).astype(f32)) / self.s).astype(f32)

    def inverse_transform_vector(self, v):
        return (quat_rotate(quat_conj(self.r), v) / self.s).astype(f32)


class Instance:
    __slots__ = ("shape", "transform")

    def __init__(self, shape, transform):
        self.shape, self.transform = shape, transform  # shape: None | ("sphere", r, cx, cy, cz) | ...


# ------------------------------------------------------------------------------------------------
# parameters (meta/params.rs)

def _src_eval(src, values):
    if src.tag == "Fixed":
        return src.fields
    off, scale = src["mapping"]["offset"], src["mapping"]["scale"]
    return f32(off) + f32(scale) * f32(values[src["idx"]])


def _src_dep(src):
    return [] if src.tag == "Fixed" else [src["idx"]]


def _spec_deps(spec):
    if spec.tag == "Constant":
        return _src_dep(spec.fields)
    return [d for k in spec.fields.values() for d in _src_dep(k)]


def _sample_spec(spec, values, rng, discrete):
    if discrete:
        if spec.tag == "Constant":
            s = spec.fields
            if s.tag == "Fixed":
                return int(s.fields)
            return int(max(np.round(_src_eval(s, values)), 0.0))

        def ev(s):
            return int(s.fields) if s.tag == "Fixed" else int(max(np.round(_src_eval(s, values)), 0.0))

        lo = ev(spec["min"])
        hi = max(ev(spec["max"]), lo)
        return rng.u32_inclusive(lo, hi)
    if spec.tag == "Constant":
        return f32(_src_eval(spec.fields, values))
    if spec.tag == "Uniform":
        lo = f32(_src_eval(spec["min"], values))
        hi = max(f32(_src_eval(spec["max"], values)), lo)
        return rng.f32_in_range(lo, hi)
    if spec.tag == "UniformCosAngle":
        d2r = f32(math.pi) / f32(180.0)
        lo = f32(_src_eval(spec["min_angle"], values)) * d2r
        hi = f32(_src_eval(spec["max_angle"], values)) * d2r
        lo = f32(min(max(lo, f32(0.0)), f32(math.pi)))
        hi = f32(min(max(hi, lo), f32(math.pi)))
        min_cos, max_cos = f32(math.cos(float(hi))), f32(math.cos(float(lo)))
        c = rng.f32_in_range(min_cos, max_cos)
        return f32(math.acos(float(c))) * (f32(180.0) / f32(math.pi))
    if spec.tag == "PowerLaw":
        lo = f32(_src_eval(spec["min"], values))
        hi = max(f32(_src_eval(spec["max"], values)), lo)
        ex = f32(_src_eval(spec["exponent"], values))
        frac = rng.f32()
        a = f32(1.0) - ex
        if abs(a) <= np.finfo(f32).eps:
            return f32(lo * f32(math.pow(float(hi / lo), float(frac))))
        lp, hp = f32(math.pow(float(lo), float(a))), f32(math.pow(float(hi), float(a)))
        return f32(math.pow(float(lp + frac * (hp - lp)), float(f32(1.0) / a)))
    raise ValueError(spec.tag)


def sample_params(node, names, rng, discrete=()):
    """`evaluate_params_for_node` (params.rs:246-264): topological order, FIFO among ready parameters."""
    specs = [node[n] for n in names]
    n = len(specs)
    dep_counts = [0] * n
    rev = [[] for _ in range(n)]
    for i, s in enumerate(specs):
        for d in _spec_deps(s):
            if d >= n:
                raise ValueError(f"Parameter {i} depends on out-of-range parameter {d}")
            dep_counts[i] += 1
            rev[d].append(i)
    queue = [i for i in range(n) if dep_counts[i] == 0]
    values = [f32(0.0)] * n
    done = 0
    while queue:
        i = queue.pop(0)
        values[i] = f32(_sample_spec(specs[i], values, rng, names[i] in discrete))
        done += 1
        for r in rev[i]:
            dep_counts[r] -= 1
            if dep_counts[r] == 0:
                queue.append(r)
    if done != n:
        raise ValueError("Cycle in parameter dependencies")
    return dict(zip(names, values))


# ------------------------------------------------------------------------------------------------
LEAF_TAGS = {"Points": 0x00, "Spheres": 0x01, "Capsules": 0x02, "Boxes": 0x03}
SEEDED_UNARY = {"Translation": 0x10, "Rotation": 0x11, "Scaling": 0x12, "Similarity": 0x13,
                "StratifiedGridTransforms": 0x14, "SphereSurfaceTransforms": 0x15, "StochasticSelection": 0x30,
                "MultifractalNoiseSDFModifier": 0x50}
UNARY = {"SDFInstantiation": 0x40, "SDFGroupUnion": 0x63}
BINARY = {"ClosestTranslationToSurface": (0x20, "surface_sdf_id", "subject_id"),
          "RayTranslationToSurface": (0x21, "surface_sdf_id", "subject_id"),
          "RotationToGradient": (0x22, "gradient_sdf_id", "subject_id"),
          "TransformApplication": (0x41, "sdf_id", "instance_id"), "SDFSubtraction": (0x61, "child_1_id", "child_2_id")}
BINARY_COMM = {"SDFUnion": 0x60, "SDFIntersection": 0x62}


def _children(node):
    t = node.tag
    if t in LEAF_TAGS:
        return []
    if t in SEEDED_UNARY or t in UNARY:
        return [node["child_id"]]
    if t in BINARY:
        return [node[BINARY[t][1]], node[BINARY[t][2]]]
    return [node["child_1_id"], node["child_2_id"]]


def _stable_seed(node, seeds):
    t = node.tag
    if t == "Points":
        return splitmix(0x00)
    if t in LEAF_TAGS:
        return splitmix2(LEAF_TAGS[t], node["seed"])
    if t in SEEDED_UNARY:
        return splitmix3(SEEDED_UNARY[t], node["seed"], seeds[node["child_id"]])
    if t in UNARY:
        return splitmix2(UNARY[t], seeds[node["child_id"]])
    if t in BINARY:
        tag, a, b = BINARY[t]
        return splitmix3(tag, seeds[node[a]], seeds[node[b]])
    s1, s2 = seeds[node["child_1_id"]], seeds[node["child_2_id"]]
    return splitmix3(BINARY_COMM[t], min(s1, s2), max(s1, s2))


def _tilt_turn_roll(tilt_deg, turn_deg, roll_deg):
    # unit_quaternion_from_tilt_turn_roll (meta.rs:2810-2834)
    d2r = f32(math.pi) / f32(180.0)
    polar, azim, roll = f32(tilt_deg) * d2r, f32(turn_deg) * d2r, f32(roll_deg) * d2r
    sp, cp = f32(math.sin(float(polar))), f32(math.cos(float(polar)))
    sa, ca = f32(math.sin(float(azim))), f32(math.cos(float(azim)))
    direction = v3(sp * ca, cp, sp * sa)
    without_roll = quat_from_rotation_arc(v3(0, 1, 0), direction)
    return quat_mul(quat_from_axis_angle(direction, roll), without_roll)


def _radial_directions(n):
    # compute_uniformly_distributed_radial_directions (impact_geometry/src/lib.rs:59-87)
    idx_norm = f32(1.0) / (f32(n - 1) if n > 1 else f32(1.0))
    golden = f32(math.pi) * (f32(3.0) - f32(np.sqrt(f32(5.0))))
    out = []
    for i in range(n):
        fi = f32(i)
        z = f32(1.0) - f32(2.0) * fi * idx_norm
        hr = f32(np.sqrt(f32(1.0) - z * z))
        az = fi * golden
        s, c = f32(math.sin(float(az))), f32(math.cos(float(az)))
        v = v3(hr * c, hr * s, z)
        out.append((v / norm(v)).astype(f32))
    return out


def _jittered_direction(direction, max_angle, rng):
    # compute_jittered_direction (meta.rs:2772-2808)
    if abs(max_angle) <= np.finfo(f32).eps:
        return direction
    angle = rng.f32_in_range(0.0, max_angle)
    axis = v3(rng.f32_in_range(-1.0, 1.0), rng.f32_in_range(-1.0, 1.0), rng.f32_in_range(-1.0, 1.0))
    axis = (axis - dot(axis, direction) * direction).astype(f32)
    n2 = dot(axis, axis)
    if n2 > f32(1e-8) * f32(1e-8):
        axis = (axis / f32(np.sqrt(n2))).astype(f32)
    else:
        axis = v3(0, 0, 1) if abs(direction[2]) < 0.9 else v3(1, 0, 0)
        axis = (axis - dot(axis, direction) * direction).astype(f32)
        axis = (axis / norm(axis)).astype(f32)
    return quat_rotate(quat_from_axis_angle(axis, angle), direction)


class MetaCompiler:
    """One `MetaSDFGraph::build_in` run, in Python: the readable mirror that tests/test_meta_native.py holds the library's
    `ivx_meta_compile` (csrc/meta.cpp, the product path: `compile_meta_nodes` below) against, node for node. `ctx`
    (impact_b200.voxel.Context) is needed only by the surface-probing nodes; graphs without them compile without a
    device."""

    def __init__(self, nodes, scale_factor=1.0, seed=0, ctx=None):
        self.nodes, self.scale, self.seed, self.ctx = nodes, f32(scale_factor), seed, ctx
        self.graph = SDFGraph()

    def build(self) -> SDFGraph:
        nodes = self.nodes
        if not nodes:
            return self.graph
        n = len(nodes)
        outputs = [("sdf", None)] * n
        state = [0] * n
        seeds = [0] * n
        stack = [("visit", n - 1)]  # root = last meta node (meta.rs:766)
        while stack:
            op, idx = stack.pop()
            if op == "visit":
                if idx >= n:
                    raise ValueError(f"Missing meta SDF node {idx}")
                if state[idx] == 2:
                    continue
                if state[idx] == 1:
                    raise ValueError("Detected cycle in meta SDF node graph")
                state[idx] = 1
                stack.append(("process", idx))
                for c in reversed(_children(nodes[idx])):
                    stack.append(("visit", c))
            else:
                state[idx] = 2
                seeds[idx] = _stable_seed(nodes[idx], seeds)
                outputs[idx] = self._resolve(nodes[idx], outputs, splitmix2(self.seed, seeds[idx]))
        kind, val = outputs[n - 1]
        if kind != "sdf":
            raise ValueError("Root meta node must have single SDF output")
        if val is None:
            return SDFGraph()
        self.graph.set_root_node(val)
        return self.graph

    # -- resolvers --------------------------------------------------------------------------------
    def _instances(self, out, name):
        if out[0] != "instances":
            raise ValueError(f"{name} node expects Instances as input, got {out[0]}")
        return out[1]

    def _per_instance(self, node, outputs, seed, names, make):
        inst = self._instances(outputs[node["child_id"]], node.tag)
        rng = Rng(seed)
        per = node["sampling"].tag == "PerInstance"
        params = sample_params(node, names, rng)
        res = []
        for i, ins in enumerate(inst):
            res.append(make(params, ins))
            if per and i + 1 < len(inst):
                params = sample_params(node, names, rng)
        return ("instances", res)

    def _resolve(self, node, outputs, seed):
        t, S, g = node.tag, self.scale, self.graph
        if t == "Points":
            return ("instances", [Instance(None, Sim()) for _ in range(node["count"])])
        if t in ("Spheres", "Capsules", "Boxes"):
            names = {"Spheres": ["radius", "center_x", "center_y", "center_z"],
                     "Capsules": ["segment_length", "radius", "center_x", "center_y", "center_z"],
                     "Boxes": ["extent_x", "extent_y", "extent_z", "center_x", "center_y", "center_z"]}[t]
            rng = Rng(seed)
            per = node["sampling"].tag == "PerInstance"
            p = sample_params(node, names, rng)
            res = []
            for i in range(node["count"]):
                res.append(Instance((t,) + tuple(f32(p[k]) * S for k in names), Sim()))
                if per and i + 1 < node["count"]:
                    p = sample_params(node, names, rng)
            return ("instances", res)
        if t == "Translation":
            def make(p, ins):
                tr = v3(p["translation_x"] * S, p["translation_y"] * S, p["translation_z"] * S)
                tf = ins.transform.translated(tr) if node["composition"].tag == "Post" else \
                    ins.transform.applied_to_translation(tr)
                return Instance(ins.shape, tf)
            return self._per_instance(node, outputs, seed, ["translation_x", "translation_y", "translation_z"], make)
        if t == "Rotation":
            def make(p, ins):
                q = _tilt_turn_roll(p["tilt_angle"], p["turn_angle"], p["roll_angle"])
                tf = ins.transform.rotated(q) if node["composition"].tag == "Post" else ins.transform.applied_to_rotation(q)
                return Instance(ins.shape, tf)
            return self._per_instance(node, outputs, seed, ["tilt_angle", "turn_angle", "roll_angle"], make)
        if t == "Scaling":
            def make(p, ins):
                s = max(f32(p["scaling"]), np.finfo(f32).eps)
                tf = ins.transform.scaled(s) if node["composition"].tag == "Post" else ins.transform.applied_to_scaling(s)
                return Instance(ins.shape, tf)
            return self._per_instance(node, outputs, seed, ["scaling"], make)
        if t == "StratifiedGridTransforms":
            inst = self._instances(outputs[node["child_id"]], t)
            if not inst:
                return ("instances", [])
            rng = Rng(seed)
            names = ["shape_x", "shape_y", "shape_z", "cell_extent_x", "cell_extent_y", "cell_extent_z", "jitter_fraction"]
            p = sample_params(node, names, rng, discrete=("shape_x", "shape_y", "shape_z"))
            shape = [int(p["shape_x"]), int(p["shape_y"]), int(p["shape_z"])]
            ext = [max(f32(p[k]) * S, f32(0.0)) for k in ("cell_extent_x", "cell_extent_y", "cell_extent_z")]
            jf = f32(min(max(p["jitter_fraction"], f32(0.0)), f32(1.0)))
            cells = shape[0] * shape[1] * shape[2]
            if cells == 0:
                return ("instances", list(inst))
            start = [f32(-0.5) * (f32(shape[d]) * ext[d]) + f32(0.5) * ext[d] for d in range(3)]
            res = []
            for idx, ins in enumerate(inst):
                c = (idx * cells) // len(inst)
                ijk = [c // (shape[1] * shape[2]), (c // shape[2]) % shape[1], c % shape[2]]
                pos = [start[d] + f32(ijk[d]) * ext[d] for d in range(3)]
                jit = [rng.f32_in_range(-0.5, 0.5) * jf * ext[d] for d in range(3)]
                tf = Sim(v3(pos[0] + jit[0], pos[1] + jit[1], pos[2] + jit[2]))
                res.append(Instance(ins.shape, tf.mul(ins.transform)))
            return ("instances", res)
        if t == "SphereSurfaceTransforms":
            inst = self._instances(outputs[node["child_id"]], t)
            if not inst:
                return ("instances", [])
            rng = Rng(seed)
            p = sample_params(node, ["radius", "jitter_fraction"], rng)
            radius = max(f32(p["radius"]) * S, f32(0.0))
            jf = f32(min(max(p["jitter_fraction"], f32(0.0)), f32(1.0)))
            count = len(inst)
            solid = f32(4.0) * f32(math.pi) / f32(count)
            x = f32(min(max(f32(1.0) - solid / (f32(2.0) * f32(math.pi)), f32(-1.0)), f32(1.0)))
            max_angle = f32(min(max(jf * f32(math.acos(float(x))), f32(0.0)), f32(0.5) * f32(math.pi)))
            res = []
            for direction, ins in zip(_radial_directions(count), inst):
                jd = _jittered_direction(direction, max_angle, rng)
                rot = {"Identity": lambda: QID.copy(), "RadialOutwards": lambda: quat_from_rotation_arc(v3(0, 1, 0), jd),
                       "RadialInwards": lambda: quat_from_rotation_arc(v3(0, -1, 0), jd)}[node["rotation"].tag]()
                res.append(Instance(ins.shape, Sim(radius * jd, rot, 1.0).mul(ins.transform)))
            return ("instances", res)
        if t == "StochasticSelection":
            rng = Rng(seed)
            lo = node["min_pick_count"]
            hi = max(node["max_pick_count"], lo)
            prob = f32(min(max(node["pick_probability"], 0.0), 1.0))
            kind, val = outputs[node["child_id"]]
            if kind == "sdf":
                if val is None:
                    return ("sdf", None)
                return ("sdf", val if (lo > 0 and rng.f32() < prob) else None)
            count = rng.u32_inclusive(lo, hi)
            sel = rng.clone_random_subset(min(count, len(val)), val)
            return (kind, [x for x in sel if rng.f32() < prob])
        if t == "SDFInstantiation":
            inst = self._instances(outputs[node["child_id"]], t)
            ids = []
            eps = np.finfo(f32).eps
            for ins in inst:
                if ins.shape is None:
                    continue
                kind = ins.shape[0]
                if kind == "Spheres":
                    nid, center = g.sphere(ins.shape[1]), ins.shape[2:5]
                elif kind == "Capsules":
                    nid, center = g.capsule(ins.shape[1], ins.shape[2]), ins.shape[3:6]
                else:
                    nid, center = g.box(list(ins.shape[1:4])), ins.shape[4:7]
                tf = ins.transform
                if any(abs(c) > eps for c in center):
                    nid = g.translation(nid, list(center))
                if abs(tf.s - f32(1.0)) > eps:
                    nid = g.scaling(nid, tf.s)
                if np.any(np.abs(tf.r - QID) > eps):
                    nid = g.rotation(nid, list(tf.r))
                if np.any(np.abs(tf.t) > eps):
                    nid = g.translation(nid, list(tf.t))
                ids.append(nid)
            return ("group", ids)
        if t == "MultifractalNoiseSDFModifier":
            kind, val = outputs[node["child_id"]]
            names = ["octaves", "frequency", "lacunarity", "persistence", "amplitude"]

            def draw(rng):
                p = sample_params(node, names, rng, discrete=("octaves",))
                return p, rng.u32_inclusive(0, 0xFFFFFFFF)

            def make(ps, child):
                p, nseed = ps
                return g.multifractal_noise(child, int(p["octaves"]), f32(p["frequency"]) / S, p["lacunarity"],
                                            p["persistence"], f32(p["amplitude"]) * S, nseed)
            if kind == "sdf":
                if val is None:
                    return ("sdf", None)
                return ("sdf", make(draw(Rng(seed)), val))
            if kind == "group":
                rng = Rng(seed)
                per = node["sampling"].tag == "PerInstance"
                ps = draw(rng)
                res = []
                for i, child in enumerate(val):
                    res.append(make(ps, child))
                    if per and i + 1 < len(val):
                        ps = draw(rng)
                return ("group", res)
            raise ValueError("MultifractalNoiseSDFModifier node expects SingleSDF or SDFGroup input, got Instances")
        if t in ("SDFUnion", "SDFSubtraction", "SDFIntersection"):
            (k1, a), (k2, b) = outputs[node["child_1_id"]], outputs[node["child_2_id"]]
            if k1 != "sdf" or k2 != "sdf":
                raise ValueError(f"{t} node expects two SingleSDF inputs, got {k1} and {k2}")
            k = max(f32(node["smoothness"]) * S, f32(0.0))
            if t == "SDFUnion":
                if a is None or b is None:
                    return ("sdf", a if b is None else b)
                return ("sdf", g.union(a, b, k))
            if t == "SDFSubtraction":
                if a is None:
                    return ("sdf", None)
                return ("sdf", a if b is None else g.subtraction(a, b, k))
            if a is None or b is None:
                return ("sdf", None)
            return ("sdf", g.intersection(a, b, k))
        if t == "SDFGroupUnion":
            kind, val = outputs[node["child_id"]]
            if kind == "sdf":
                return ("sdf", val)
            if kind != "group":
                raise ValueError("SDFGroupUnion node expects SDFGroup or SingleSDF input, got Instances")
            k = max(f32(node["smoothness"]) * S, f32(0.0))
            queue = list(val)  # emit_balanced_binary_tree (meta.rs:2390-2409)
            while len(queue) > 1:
                a, b = queue.pop(0), queue.pop(0)
                queue.append(g.union(a, b, k))
            return ("sdf", queue[0] if queue else None)
        if t == "RayTranslationToSurface":
            return self._ray_translation(node, outputs)
        if t == "Similarity":  # meta.rs:1402-1446
            def make(p, ins):
                scaling = max(f32(p["scale"]), np.finfo(f32).eps)
                q = _tilt_turn_roll(p["tilt_angle"], p["turn_angle"], p["roll_angle"])
                tr = v3(p["translation_x"] * S, p["translation_y"] * S, p["translation_z"] * S)
                tf = Sim(tr, q, scaling)
                return Instance(ins.shape, tf.mul(ins.transform) if node["composition"].tag == "Post" else ins.transform.mul(tf))
            return self._per_instance(node, outputs, seed, ["scale", "tilt_angle", "turn_angle", "roll_angle", "translation_x",
                                                            "translation_y", "translation_z"], make)
        if t == "ClosestTranslationToSurface":
            return self._closest_translation(node, outputs)
        if t == "RotationToGradient":
            return self._rotation_to_gradient(node, outputs)
        if t == "TransformApplication":  # meta.rs:2012-2075
            kind, val = outputs[node["sdf_id"]]
            if kind == "instances":
                raise ValueError("TransformApplication node expects SingleSDF or GroupSDF as input 1, got Instances")
            sdf_ids = ([] if val is None else [val]) if kind == "sdf" else list(val)
            k2, inst = outputs[node["instance_id"]]
            if k2 != "instances":
                raise ValueError(f"TransformApplication node expects Instances as input 2, got {k2}")
            eps = np.finfo(f32).eps
            ids = []
            for sid in sdf_ids:
                for ins in inst:
                    tf, nid = ins.transform, sid
                    if abs(tf.s - f32(1.0)) > eps:
                        nid = g.scaling(nid, tf.s)
                    if np.any(np.abs(tf.r - QID) > eps):
                        nid = g.rotation(nid, list(tf.r))
                    if np.any(np.abs(tf.t) > eps):
                        nid = g.translation(nid, list(tf.t))
                    ids.append(nid)
            return ("group", ids)
        raise ValueError(f"unknown meta node kind {t}")

    # -- the two nodes that sample an SDF's value and gradient around the subject's centre (meta.rs:1620-1688, 1798-1862,
    #    2411-2540, 2728-2769): 2x2x2 blocks, all instances in lock step through ivx_program_eval_blocks ------------------
    def _surface_generator(self, sdf_id, what):
        from .voxel import SDFGenerator

        if self.ctx is None:
            raise RuntimeError(f"{what} needs a device context for its SDF probes")
        nodes = self.graph.nodes()
        gen = SDFGenerator.from_graph(self.ctx, nodes, sdf_id)
        sn = nodes[sdf_id]  # node_to_parent_transform of the sampled node (atomic.rs:1138-1148)
        if sn["kind"] == 3:
            surf = Sim(t=sn["p"][:3])
        elif sn["kind"] == 4:
            surf = Sim(r=sn["p"][:4])
        elif sn["kind"] == 5:
            surf = Sim(s=sn["p"][0])
        else:
            surf = Sim()
        return gen, surf

    @staticmethod
    def _sample_with_gradient(gen, pos):
        """sample_signed_distance_with_gradient for many positions → (centre values, gradients)."""
        d = gen.compute_signed_distances_for_blocks_preserving_gradients((pos - f32(0.5)).astype(f32), 2)
        total = np.zeros(len(pos), f32)
        for q in range(8):  # iter().sum::<f32>() in sample order
            total = (total + d[:, q]).astype(f32)
        d000, d001, d010, d011, d100, d101, d110, d111 = [d[:, q] for q in range(8)]
        grad = (f32(0.25) * np.stack([
            (d100 + d110 + d101 + d111) - (d000 + d010 + d001 + d011),
            (d010 + d110 + d011 + d111) - (d000 + d100 + d001 + d101),
            (d001 + d101 + d011 + d111) - (d000 + d100 + d010 + d110)], 1)).astype(f32)
        return (total * f32(0.125)).astype(f32), grad

    def _closest_translation(self, node, outputs):
        subjects = self._instances(outputs[node["subject_id"]], "ClosestTranslationToSurface")
        kind, sdf_id = outputs[node["surface_sdf_id"]]
        if kind != "sdf":
            raise ValueError(f"ClosestTranslationToSurface node expects SingleSDF as input 1, got {kind}")
        if sdf_id is None or not subjects:
            return ("instances", list(subjects))
        gen, surf = self._surface_generator(sdf_id, "ClosestTranslationToSurface")
        n = len(subjects)
        start = np.array([surf.inverse_transform_point(ins.transform.transform_point(v3(0, 0, 0))) for ins in subjects],
                         f32).reshape(n, 3)
        pos = start.copy()
        alive = np.ones(n, bool)
        active = np.ones(n, bool)
        for _ in range(5):  # Newton-Raphson, max_iterations = 5, max_distance_from_surface = 0.1
            idx = np.flatnonzero(active)
            if len(idx) == 0:
                break
            sd, grad = self._sample_with_gradient(gen, pos[idx])
            n2 = ((grad[:, 0] * grad[:, 0] + grad[:, 1] * grad[:, 1]) + grad[:, 2] * grad[:, 2]).astype(f32)
            flat = np.abs(n2) <= f32(1e-8)
            alive[idx[flat]] = False
            active[idx[flat]] = False
            ok = ~flat
            step = ((-sd[ok] / n2[ok])[:, None] * grad[ok]).astype(f32)
            pos[idx[ok]] = (pos[idx[ok]] + step).astype(f32)
            active[idx[ok][np.abs(sd[ok]) <= f32(0.1)]] = False
        res = []
        for i, ins in enumerate(subjects):
            if alive[i]:
                tr = surf.transform_vector((pos[i] - start[i]).astype(f32))
                res.append(Instance(ins.shape, ins.transform.translated(tr)))
        return ("instances", res)

    def _rotation_to_gradient(self, node, outputs):
        subjects = self._instances(outputs[node["subject_id"]], "RotationToGradient")
        kind, sdf_id = outputs[node["gradient_sdf_id"]]
        if kind != "sdf":
            raise ValueError(f"RotationToGradient node expects SingleSDF as input 1, got {kind}")
        if sdf_id is None or not subjects:
            return ("instances", list(subjects))
        gen, surf = self._surface_generator(sdf_id, "RotationToGradient")
        n = len(subjects)
        centre = np.array([surf.inverse_transform_point(ins.transform.transform_point(v3(0, 0, 0))) for ins in subjects],
                          f32).reshape(n, 3)
        _, grad = self._sample_with_gradient(gen, centre)
        res = []
        tiny = f32(1e-8) * f32(1e-8)
        for i, ins in enumerate(subjects):
            y_axis = ins.transform.transform_vector(v3(0, 1, 0))
            gp = surf.transform_vector(grad[i].astype(f32))
            ny, ng = dot(y_axis, y_axis), dot(gp, gp)
            if not (ny > tiny and ng > tiny):
                continue
            q = quat_from_rotation_arc((y_axis / f32(np.sqrt(ny))).astype(f32), (gp / f32(np.sqrt(ng))).astype(f32))
            res.append(Instance(ins.shape, ins.transform.rotated(q)))
        return ("instances", res)

    # -- RayTranslationToSurface (meta.rs:1690-1796, 2534-2748), all instances in lock step ------------------
    def _ray_translation(self, node, outputs):
        from .voxel import SDFGenerator

        subjects = self._instances(outputs[node["subject_id"]], "RayTranslationToSurface")
        kind, sdf_id = outputs[node["surface_sdf_id"]]
        if kind != "sdf":
            raise ValueError(f"RayTranslationToSurface node expects SingleSDF as input 1, got {kind}")
        if sdf_id is None:
            return ("instances", list(subjects))
        if self.ctx is None:
            raise RuntimeError("RayTranslationToSurface needs a device context for its surface probes")
        nodes = self.graph.nodes()
        gen = SDFGenerator.from_graph(self.ctx, nodes, sdf_id)
        dom_lo, dom_hi = gen.domain
        # node_to_parent_transform of the surface node (atomic.rs:1138-1148)
        sn = nodes[sdf_id]
        if sn["kind"] == 3:
            surf = Sim(t=sn["p"][:3])
        elif sn["kind"] == 4:
            surf = Sim(r=sn["p"][:4])
        elif sn["kind"] == 5:
            surf = Sim(s=sn["p"][0])
        else:
            surf = Sim()
        anchor_shape = node["anchor"].tag == "ShapeBoundaryAtOrigin"

        # per instance: sphere + direction in the surface node's space
        origins, dirs, radii, alive = [], [], [], []
        for ins in subjects:
            c, r = v3(0, 0, 0), f32(0.0)
            if anchor_shape and ins.shape is not None:
                sh = ins.shape
                if sh[0] == "Spheres":
                    r = f32(sh[1])
                elif sh[0] == "Capsules":
                    c, r = v3(0, f32(0.5) * f32(sh[1]), 0), f32(sh[2])
                else:
                    r = f32(0.5) * min(f32(sh[1]), f32(sh[2]), f32(sh[3]))
                    c = v3(0, f32(0.5) * f32(sh[2]) - r, 0)
            tf = ins.transform
            cp = tf.transform_point(c)
            rp = tf.s * r
            dp = tf.transform_vector(v3(0, 1, 0))
            cs = surf.inverse_transform_point(cp)
            rs = (f32(1.0) / surf.s) * rp
            ds = surf.inverse_transform_vector(dp)
            n2 = dot(ds, ds)
            ok = n2 > f32(1e-8) * f32(1e-8)
            origins.append(cs)
            radii.append(rs)
            dirs.append((ds / f32(np.sqrt(n2))).astype(f32) if ok else v3(0, 1, 0))
            alive.append(bool(ok))
        n = len(subjects)
        origins, dirs = np.array(origins, f32).reshape(n, 3), np.array(dirs, f32).reshape(n, 3)
        radii, alive = np.array(radii, f32), np.array(alive, bool)

        # ray / domain intersection (axis_aligned_box.rs:420-455)
        t_start, t_end = np.zeros(n, f32), np.zeros(n, f32)
        for i in range(n):
            if not alive[i]:
                continue
            tmin, tmax, hit = f32(0.0), f32(np.inf), True
            for d in range(3):
                if dirs[i, d] != 0.0:
                    rc = f32(1.0) / dirs[i, d]
                    t1, t2 = (dom_lo[d] - origins[i, d]) * rc, (dom_hi[d] - origins[i, d]) * rc
                    te, tx = (t1, t2) if t1 < t2 else (t2, t1)
                    tmin, tmax = max(tmin, te), min(tmax, tx)
                    if tmax < tmin:
                        hit = False
                        break
                elif origins[i, d] < dom_lo[d] or origins[i, d] > dom_hi[d]:
                    hit = False
                    break
            if not hit or tmax < 0.0:
                alive[i] = False
                continue
            t_start[i], t_end[i] = max(tmin, f32(0.0)) - radii[i], tmax

        def smallest_sd(pos, mask):
            """compute_smallest_signed_distance_on_sphere for the masked instances → (values, ok)"""
            idx = np.flatnonzero(mask)
            val = np.zeros(n, f32)
            ok = mask.copy()
            if len(idx) == 0:
                return val, ok
            probe = pos[idx].copy()
            with_r = np.abs(radii[idx]) > np.finfo(f32).eps
            if with_r.any():
                blk = (pos[idx[with_r]] - f32(0.5)).astype(f32)  # 2x2x2 block around the position
                d = gen.compute_signed_distances_for_blocks_preserving_gradients(blk, 2)
                d000, d001, d010, d011, d100, d101, d110, d111 = [d[:, q] for q in range(8)]
                grad = f32(0.25) * np.stack([
                    (d100 + d110 + d101 + d111) - (d000 + d010 + d001 + d011),
                    (d010 + d110 + d011 + d111) - (d000 + d100 + d001 + d101),
                    (d001 + d101 + d011 + d111) - (d000 + d100 + d010 + d110)], 1).astype(f32)
                n2 = ((grad[:, 0] * grad[:, 0] + grad[:, 1] * grad[:, 1]) + grad[:, 2] * grad[:, 2]).astype(f32)
                good = n2 > f32(1e-8) * f32(1e-8)
                gdir = grad / np.sqrt(np.where(good, n2, f32(1.0)))[:, None]
                sub = np.flatnonzero(with_r)
                probe[sub] = (pos[idx[with_r]] - radii[idx[with_r], None] * gdir).astype(f32)
                ok[idx[sub[~good]]] = False
            v = gen.compute_signed_distances_for_blocks_preserving_gradients(probe, 1)[:, 0]
            val[idx] = v
            return val, ok

        dist = t_start.copy()
        pos = (origins + dist[:, None] * dirs).astype(f32)
        sd, ok = smallest_sd(pos, alive)
        alive &= ok
        alive &= ~(sd < 0.0)  # already penetrating: a miss (meta.rs:2646-2650)
        active = alive & (np.abs(sd) > f32(0.1))
        crossed = np.zeros(n, bool)
        step = 0
        while active.any():
            step += 1
            if step >= 128:
                alive &= ~(active & ~crossed)  # gave up without crossing: a miss
                break
            dist = np.where(active, dist + sd * f32(0.5), dist).astype(f32)
            crossed |= active & np.signbit(sd)
            out = active & ((dist > t_end) | (dist < t_start))
            alive &= ~out
            active &= ~out
            pos = np.where(active[:, None], origins + dist[:, None] * dirs, pos).astype(f32)
            nsd, ok = smallest_sd(pos, active)
            alive &= ~(active & ~ok)
            active &= ok
            sd = np.where(active, nsd, sd).astype(f32)
            active &= np.abs(sd) > f32(0.1)
        res = []
        for i, ins in enumerate(subjects):
            if not alive[i]:
                continue
            tr_surface = (pos[i] - origins[i]).astype(f32)
            tr_parent = surf.transform_vector(tr_surface)
            res.append(Instance(ins.shape, ins.transform.translated(tr_parent)))
        return ("instances", res)


# ------------------------------------------------------------------------------------------------
# the product path: the same compile inside the library (csrc/meta.cpp) behind `ivx_meta_compile`. The class above is the
# test mirror it is compared with node for node (tests/test_meta.py); hosts other than Python hand the library PODs.

META_KIND_IDS = {k: i for i, k in enumerate(EDITOR_NODE_KINDS)}
# the kind's distributed parameters in the reference struct's declaration order (= the index space of `FromParam`)
META_PARAM_NAMES = {
    "Spheres": ["radius", "center_x", "center_y", "center_z"],
    "Capsules": ["segment_length", "radius", "center_x", "center_y", "center_z"],
    "Boxes": ["extent_x", "extent_y", "extent_z", "center_x", "center_y", "center_z"],
    "Translation": ["translation_x", "translation_y", "translation_z"],
    "Rotation": ["tilt_angle", "turn_angle", "roll_angle"],
    "Scaling": ["scaling"],
    "Similarity": ["scale", "tilt_angle", "turn_angle", "roll_angle", "translation_x", "translation_y", "translation_z"],
    "StratifiedGridTransforms": ["shape_x", "shape_y", "shape_z", "cell_extent_x", "cell_extent_y", "cell_extent_z",
                                 "jitter_fraction"],
    "SphereSurfaceTransforms": ["radius", "jitter_fraction"],
    "MultifractalNoiseSDFModifier": ["octaves", "frequency", "lacunarity", "persistence", "amplitude"],
}
_DIST_IDS = {"Constant": 0, "Uniform": 1, "UniformCosAngle": 2, "PowerLaw": 3}
_DIST_FIELDS = {"Uniform": ("min", "max"), "UniformCosAngle": ("min_angle", "max_angle"), "PowerLaw": ("min", "max", "exponent")}


def meta_nodes_to_pod(nodes):
    """Meta nodes (as `load_vgen_ron` / `load_graph_ron` / `asteroid_meta_nodes` give them) → ctypes array of
    `ivx_meta_node`."""
    from . import _lib as L

    def put_source(dst, src):
        if src.tag == "Fixed":
            dst.kind, dst.idx, dst.value, dst.scale = 0, 0, float(src.fields), 0.0
        else:
            dst.kind, dst.idx = 1, int(src["idx"])
            dst.value, dst.scale = float(src["mapping"]["offset"]), float(src["mapping"]["scale"])

    arr = (L.MetaNode * max(len(nodes), 1))()
    for pod, node in zip(arr, nodes):
        t = node.tag
        if t not in META_KIND_IDS:
            raise ValueError(f"unknown meta node kind {t}")
        pod.kind = META_KIND_IDS[t]
        for slot, c in enumerate(_children(node)):
            pod.child[slot] = int(c)
        f = node.fields or {}
        pod.count = int(f.get("count", 0))
        pod.seed = int(f.get("seed", 0))
        pod.sampling = _SAMPLING.index(f["sampling"].tag) if "sampling" in f else 0
        pod.composition = _COMPOSITION.index(f["composition"].tag) if "composition" in f else 0
        pod.rotation = _ROTATION.index(f["rotation"].tag) if "rotation" in f else 0
        pod.anchor = _ANCHOR.index(f["anchor"].tag) if "anchor" in f else 0
        pod.min_pick_count = int(f.get("min_pick_count", 0))
        pod.max_pick_count = int(f.get("max_pick_count", 0))
        pod.pick_probability = float(f.get("pick_probability", 0.0))
        pod.smoothness = float(f.get("smoothness", 0.0))
        for i, name in enumerate(META_PARAM_NAMES.get(t, ())):
            spec = f[name]
            pod.params[i].dist = _DIST_IDS[spec.tag]
            if spec.tag == "Constant":
                put_source(pod.params[i].src[0], spec.fields)
            else:
                for q, key in enumerate(_DIST_FIELDS[spec.tag]):
                    put_source(pod.params[i].src[q], spec[key])
    return arr


def compile_meta_nodes(nodes, scale_factor=1.0, seed=0, ctx=None) -> SDFGraph:
    """`MetaSDFGraph::build_in` through the library's `ivx_meta_compile` → atomic SDFGraph (empty when the meta graph
    resolves to nothing). `ctx` is needed only for graphs with surface-probing nodes."""
    import ctypes as C

    from . import _lib as L
    from .graph import SDF_NODE_DTYPE

    lib = L.lib()
    pods = meta_nodes_to_pod(nodes)
    handle = ctx.h if ctx is not None else None
    count, root, empty = C.c_uint32(), C.c_uint32(), C.c_int()
    err = C.create_string_buffer(512)
    capacity = 4096
    while True:
        out = np.zeros(capacity, SDF_NODE_DTYPE)
        rc = lib.ivx_meta_compile(handle, pods, C.c_uint32(len(nodes)),
                                  C.c_float(scale_factor), C.c_uint64(seed & M64), L.ptr(out), C.c_uint32(capacity),
                                  C.byref(count), C.byref(root), C.byref(empty), err, C.c_size_t(len(err)))
        if rc == 5 and count.value > capacity:  # IVX_ERR_CAPACITY
            capacity = count.value
            continue
        break
    if rc != 0:
        msg = err.value.decode() or f"ivx_meta_compile failed with status {rc}"
        raise (RuntimeError if rc == 1 else ValueError)(msg)
    if empty.value:
        return SDFGraph()
    return graph_from_nodes(out[: count.value], root.value)


def _fixed(v):
    return Tagged("Fixed", v)


def _const(v):
    return Tagged("Constant", _fixed(v))


def _from_param(idx, scale, offset=0.0):
    return Tagged("FromParam", {"idx": idx, "mapping": Tagged("Linear", {"offset": offset, "scale": scale})})


def asteroid_meta_nodes():
    """The 29 meta nodes of engine/benches/data/asteroid.vgen.ron (:1-294), transcribed as data: a body of 3-6
    smooth-unioned spheres with 1-octave noise, three crater passes (40 / 150 / 250 capsules placed on a sphere,
    ray-cast onto the current surface, group-unioned and smooth-subtracted) and a final 5-octave noise.
    tests/test_meta.py checks this transcription against the RON file whenever /root/reference is present."""
    T = Tagged
    per, once = T("PerInstance", None), T("OnlyOnce", None)
    n = [
        T("Spheres", {"radius": T("Uniform", {"min": _fixed(30.0), "max": _fixed(60.0)}), "center_x": _const(0.0),
                      "center_y": _const(0.0), "center_z": _const(0.0), "count": 8, "seed": 0, "sampling": per}),
        T("StratifiedGridTransforms", {"child_id": 0, "shape_x": _const(2), "shape_y": _const(2), "shape_z": _const(2),
                                       "cell_extent_x": _const(30.0), "cell_extent_y": _const(30.0),
                                       "cell_extent_z": _const(30.0), "jitter_fraction": _const(0.75), "seed": 0}),
        T("StochasticSelection", {"child_id": 1, "min_pick_count": 3, "max_pick_count": 6, "pick_probability": 1.0, "seed": 0}),
        T("Scaling", {"child_id": 2, "composition": T("Post", None),
                      "scaling": T("Uniform", {"min": _fixed(0.5), "max": _fixed(2.0)}), "seed": 0, "sampling": per}),
        T("SDFInstantiation", {"child_id": 3}),
        T("SDFGroupUnion", {"child_id": 4, "smoothness": 25.0}),
        T("MultifractalNoiseSDFModifier", {"child_id": 5, "octaves": _const(1), "frequency": _const(0.01),
                                           "lacunarity": _const(2.0), "persistence": _const(0.5), "amplitude": _const(8.0),
                                           "seed": 0, "sampling": once}),
    ]
    surface = 6
    passes = [((50.0, 80.0), (0.2, 0.3), 40, 350.0, 8.0), ((15.0, 55.0), (0.15, 0.3), 150, 300.0, 3.0),
              ((5.0, 25.0), (0.25, 0.4), 250, 250.0, 3.0)]
    for (rmin, rmax), (ylo, yhi), count, shell, k in passes:
        b = len(n)
        n += [
            T("Capsules", {"segment_length": T("Constant", _from_param(1, 1.0)),
                           "radius": T("PowerLaw", {"min": _fixed(rmin), "max": _fixed(rmax), "exponent": _fixed(2.0)}),
                           "center_x": _const(0.0),
                           "center_y": T("Uniform", {"min": _from_param(1, ylo), "max": _from_param(1, yhi)}),
                           "center_z": _const(0.0), "count": count, "seed": 0, "sampling": per}),
            T("Rotation", {"child_id": b, "composition": T("Post", None),
                           "tilt_angle": T("UniformCosAngle", {"min_angle": _fixed(0.0), "max_angle": _fixed(80.0)}),
                           "turn_angle": T("UniformCosAngle", {"min_angle": _fixed(0.0), "max_angle": _fixed(360.0)}),
                           "roll_angle": _const(0.0), "seed": 0, "sampling": per}),
            T("SphereSurfaceTransforms", {"child_id": b + 1, "radius": _const(shell), "jitter_fraction": _const(1.0),
                                          "rotation": T("RadialInwards", None), "seed": 0}),
            T("RayTranslationToSurface", {"surface_sdf_id": surface, "subject_id": b + 2,
                                          "anchor": T("ShapeBoundaryAtOrigin", None)}),
            T("SDFInstantiation", {"child_id": b + 3}),
            T("SDFGroupUnion", {"child_id": b + 4, "smoothness": k}),
            T("SDFSubtraction", {"child_1_id": surface, "child_2_id": b + 5, "smoothness": k}),
        ]
        surface = b + 6
    n.append(T("MultifractalNoiseSDFModifier", {"child_id": surface, "octaves": _const(5), "frequency": _const(0.02),
                                                "lacunarity": _const(2.0), "persistence": _const(0.546),
                                                "amplitude": _const(2.0), "seed": 0, "sampling": once}))
    return n


def compile_file(path, scale_factor=1.0, seed=0, ctx=None) -> SDFGraph:
    return compile_meta_nodes(load_vgen_ron(path), scale_factor, seed, ctx)


_CACHE = {}
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def graph_from_nodes(nodes: np.ndarray, root: int) -> SDFGraph:
    g = SDFGraph()
    for n in nodes:
        g.add_node((int(n["kind"]), (int(n["child"][0]), int(n["child"][1])), int(n["octaves"]), int(n["seed"]),
                    [np.float32(x) for x in n["p"]]))
    g.set_root_node(int(root))
    return g


def asteroid_graph_scaled(max_dim_lo, max_dim_hi, seed=0, ctx=None, use_cache=True):
    """BASELINE configs 3-5: the asteroid meta graph compiled with `seed`, `scale_factor` tuned so the largest
    grid dimension lies in (max_dim_lo, max_dim_hi]. The compile needs a device (surface probes); the compiled
    atomic graphs of the bench workloads are kept under impact_b200/data/ (written by
    `python -m impact_b200.meta`, re-checked against a fresh compile by tests/test_meta.py) so that every arm of
    bench.py — including the CPU reference arm — starts from the identical atomic graph."""
    from . import workloads as W
    from .voxel import Context

    key = (max_dim_lo, max_dim_hi, seed)
    if key in _CACHE:
        return _CACHE[key]
    path = os.path.join(DATA_DIR, f"asteroid_{max_dim_hi}_seed{seed}.npz")
    if use_cache and os.path.exists(path):
        z = np.load(path)
        g = graph_from_nodes(z["nodes"], int(z["root"]))
        _CACHE[key] = g
        return g
    own = ctx is None
    if own:
        ctx = Context(0)
    nodes = asteroid_meta_nodes()
    try:
        s = W.scale_to_max_dim(lambda sc: compile_meta_nodes(nodes, sc, seed, ctx), max_dim_lo, max_dim_hi,
                               max_dim_hi / 330.0)
        g = compile_meta_nodes(nodes, s, seed, ctx)
    finally:
        if own:
            ctx.close()
    g.scale_factor = float(s)
    _CACHE[key] = g
    return g


if __name__ == "__main__":  # writes the bench workloads' atomic graphs (run on a GPU box)
    import sys

    out_dir = sys.argv[1] if len(sys.argv) > 1 else DATA_DIR
    os.makedirs(out_dir, exist_ok=True)
    for hi in (128, 512, 1024, 2048):
        g = asteroid_graph_scaled(hi - 16, hi, 0, use_cache=False)
        np.savez(os.path.join(out_dir, f"asteroid_{hi}_seed0.npz"), nodes=g.nodes(), root=np.int64(g.root_node_id),
                 scale_factor=np.float64(getattr(g, "scale_factor", 0.0)))
        print(hi, "nodes", len(g), "scale", getattr(g, "scale_factor", None))
