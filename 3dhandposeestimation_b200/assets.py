"""MANO model assets: pkl reader, seeded synthetic MANO-shaped model, and the
host-side constant packer that lays the model out for the sm_100a kernels.

Reference behaviour being replaced: ``ManoLayer.__init__``
(/root/reference/network/sub_modules/MANOLayer.py:52-80) unpickles
``MANO_RIGHT.pkl`` (python-2 pickle, latin1, chumpy objects inside) and casts
eight constants to fp32.  This module reads the same file without chumpy (a
allow-listing unpickler: ``chumpy.*`` classes become inert stubs, numpy / scipy.sparse array globals resolve, anything else is refused), and can build a
synthetic model of identical keys/shapes/sparsity so that nothing
MANO-licensed has to live in this repository.

No arithmetic of the hot path happens here; only layout work done once per
module construction.
"""
from __future__ import annotations

import io
import pickle
import warnings
from dataclasses import dataclass

import numpy as np

N_VERTS = 778
N_JOINTS = 16          # kinematic-chain joints (kintree_table.shape[1])
N_OUT_JOINTS = 21      # 16 chain joints + 5 fingertip vertices
N_BETAS = 10
N_POSE_FEAT = 135      # 15 joints * 9 (R - I)
N_POSE_AA = 45         # 15 joints * 3 axis-angle
N_VC = N_VERTS * 3     # 2334 vertex coordinates
# Fingertip vertices and the output slots they are inserted at
# (MANOLayer.py:196-200).
TIP_VERTS = (333, 444, 672, 555, 745)
TIP_SLOTS = (4, 8, 12, 16, 20)
# Output slot of chain joint k (what is left of 0..20 after the tip inserts).
CHAIN_SLOTS = (0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 13, 14, 15, 17, 18, 19)

# Feature vector layout of the blend contraction  f = [beta(10) | pf(135) | 1 | 0 pad]
FEAT_K = 148           # 146 rounded up to a multiple of 4 (float4 rows)
FEAT_ONE = 145         # index of the constant-1 feature that carries v_template
MAX_INFL = 8           # ELL width of the skinning weights (real MANO: <= 6)


class _ChStub:
    """Inert stand-in for chumpy objects found inside MANO pickles."""

    def __setstate__(self, state):
        self.__dict__.update(state)


# The only globals a MANO pickle resolves (traced on MANO_RIGHT.pkl): numpy array reconstruction, a scipy CSC
# matrix (J_regressor), a builtin set, and chumpy nodes (stubbed).  Anything else is refused, so a crafted pickle
# cannot reach os.system & co. through this reader (the reference's plain pickle.load would).
_ALLOWED_GLOBALS = {
    ("numpy", "dtype"), ("numpy", "ndarray"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("scipy.sparse.csc", "csc_matrix"), ("scipy.sparse._csc", "csc_matrix"), ("scipy.sparse", "csc_matrix"),
    ("scipy.sparse.csr", "csr_matrix"), ("scipy.sparse._csr", "csr_matrix"), ("scipy.sparse", "csr_matrix"),
    ("__builtin__", "set"), ("builtins", "set"), ("__builtin__", "object"), ("builtins", "object"),
    ("copy_reg", "_reconstructor"), ("copyreg", "_reconstructor"),
}


class _ManoUnpickler(pickle.Unpickler):
    """Allow-listing unpickler: chumpy classes become inert stubs, the globals in ``_ALLOWED_GLOBALS`` resolve
    normally, everything else raises ``pickle.UnpicklingError``."""

    def find_class(self, module, name):
        if module.split(".")[0] == "chumpy":
            return type(name, (_ChStub,), {})
        if (module, name) not in _ALLOWED_GLOBALS:
            raise pickle.UnpicklingError(f"MANO pickle asks for {module}.{name}, which a MANO model never needs")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)     # scipy.sparse.csc is a deprecated alias
            return super().find_class(module, name)


def _ch_to_numpy(obj) -> np.ndarray:
    """Evaluate the two chumpy node kinds that occur in MANO pickles:
    ``Ch`` (leaf holding ``x``) and ``reordering.Select`` (``a.ravel()[idxs]``
    reshaped to ``preferred_shape``)."""
    if isinstance(obj, np.ndarray):
        return obj
    d = obj.__dict__
    if "idxs" in d and "a" in d:
        base = _ch_to_numpy(d["a"])
        return np.asarray(base).ravel()[np.asarray(d["idxs"])].reshape(d["preferred_shape"])
    if "x" in d:
        return np.asarray(d["x"])
    raise TypeError(f"unsupported chumpy node {type(obj).__name__} with keys {sorted(d)}")


def read_mano_pkl(path) -> dict:
    """Read a MANO pickle into plain float64/int numpy arrays.

    Raises FileNotFoundError like the reference does (MANOLayer.py:63)."""
    with open(path, "rb") as fh:
        raw = _ManoUnpickler(io.BytesIO(fh.read()), encoding="latin1").load()
    jreg = raw["J_regressor"]
    if hasattr(jreg, "todense"):
        jreg = np.asarray(jreg.todense())
    out = {
        "v_template": np.asarray(_ch_to_numpy(raw["v_template"]), dtype=np.float64),
        "shapedirs": np.asarray(_ch_to_numpy(raw["shapedirs"]), dtype=np.float64),
        "posedirs": np.asarray(_ch_to_numpy(raw["posedirs"]), dtype=np.float64),
        "J_regressor": np.asarray(jreg, dtype=np.float64),
        "weights": np.asarray(_ch_to_numpy(raw["weights"]), dtype=np.float64),
        "hands_components": np.asarray(raw["hands_components"], dtype=np.float64),
        "hands_mean": np.asarray(raw["hands_mean"], dtype=np.float64),
        "kintree_table": np.asarray(raw["kintree_table"]),
        "f": np.asarray(raw["f"]),
    }
    _check_shapes(out)
    return out


def _check_shapes(m: dict) -> None:
    exp = {
        "v_template": (N_VERTS, 3),
        "shapedirs": (N_VERTS, 3, N_BETAS),
        "posedirs": (N_VERTS, 3, N_POSE_FEAT),
        "J_regressor": (N_JOINTS, N_VERTS),
        "weights": (N_VERTS, N_JOINTS),
        "hands_components": (N_POSE_AA, N_POSE_AA),
        "hands_mean": (N_POSE_AA,),
        "kintree_table": (2, N_JOINTS),
    }
    for key, shape in exp.items():
        if tuple(m[key].shape) != shape:
            raise ValueError(f"MANO model field {key!r} has shape {tuple(m[key].shape)}, expected {shape}")


MANO_PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)


def synthetic_mano(seed: int = 20261018) -> dict:
    """Seeded synthetic model with MANO's keys, shapes, kinematic tree, value
    ranges and sparsity (J_regressor ~1.9k nnz, skin weights 1..6 per vertex
    summing to 1, non-unit-norm orthogonal PCA rows, non-zero pose mean).

    ``numpy.random.RandomState`` (MT19937) has a frozen stream, so the model is
    bit-identical wherever it is generated; tests pin a checksum."""
    rs = np.random.RandomState(seed)
    lo = np.array([-0.079, -0.035, -0.052])
    hi = np.array([0.114, 0.032, 0.050])
    v_template = lo + (hi - lo) * rs.rand(N_VERTS, 3)

    shapedirs = rs.randn(N_VERTS, 3, N_BETAS) * 0.011
    posedirs = rs.randn(N_VERTS, 3, N_POSE_FEAT) * 0.0004
    posedirs *= rs.rand(N_VERTS, 3, N_POSE_FEAT) < 0.37

    jreg = np.zeros((N_JOINTS, N_VERTS))
    for j in range(N_JOINTS):
        idx = rs.choice(N_VERTS, size=118 + (j % 3), replace=False)
        w = rs.rand(idx.size) + 0.05
        jreg[j, idx] = w / w.sum()

    # Skinning weights with MANO's locality: the real model's vertex order walks the surface, so
    # runs of consecutive vertices share their dominant bone (about 3 distinct dominant bones per
    # 32 vertices, per-warp max of ~4 influences, 35% of the vertices on the palm bone 0) and the
    # secondary influences are tree neighbours.  Some weights are tiny (1e-5 .. 1e-3) like MANO's.
    children = {j: [c for c in range(N_JOINTS) if MANO_PARENTS[c] == j] for j in range(N_JOINTS)}
    weights = np.zeros((N_VERTS, N_JOINTS))
    v = 0
    while v < N_VERTS:
        run = int(rs.randint(4, 22))
        prim = 0 if rs.rand() < 0.35 else int(rs.randint(1, N_JOINTS))
        neigh = [b for b in [MANO_PARENTS[prim]] + children[prim] if b >= 0]
        if prim != 0:
            neigh += [b for b in children.get(MANO_PARENTS[prim], []) if b != prim][:2]
        base = int(rs.choice([1, 2, 3, 4, 5], p=[0.2, 0.15, 0.33, 0.25, 0.07]))
        for u in range(v, min(v + run, N_VERTS)):
            cnt = int(np.clip(base + rs.choice([-1, 0, 0, 1]), 1, min(6, 1 + len(neigh))))
            if rs.rand() < 0.02:
                cnt = min(6, 1 + len(neigh))
            others = list(rs.choice(neigh, size=cnt - 1, replace=False)) if cnt > 1 else []
            w = np.concatenate([[1.0 + rs.rand()], rs.rand(cnt - 1) * 0.6 + 0.01])
            tiny = rs.rand(cnt) < 0.12
            tiny[0] = False
            w = np.where(tiny, 10.0 ** rs.uniform(-5, -3, size=cnt), w)
            weights[u, [prim] + [int(b) for b in others]] = w / w.sum()
        v += run

    q, _ = np.linalg.qr(rs.randn(N_POSE_AA, N_POSE_AA))
    row_norm = 1.35 * np.exp(-np.arange(N_POSE_AA) / 18.0) + 0.05
    hands_components = q * row_norm[:, None]
    hands_mean = rs.uniform(-0.5, 0.85, size=N_POSE_AA)

    kintree = np.zeros((2, N_JOINTS), dtype=np.int64)
    kintree[0] = np.array(MANO_PARENTS, dtype=np.int64)
    kintree[0, 0] = 4294967295
    kintree[1] = np.arange(N_JOINTS)
    faces = rs.randint(0, N_VERTS, size=(1538, 3)).astype(np.uint32)
    return {
        "v_template": v_template,
        "shapedirs": shapedirs,
        "posedirs": posedirs,
        "J_regressor": jreg,
        "weights": weights,
        "hands_components": hands_components,
        "hands_mean": hands_mean,
        "kintree_table": kintree,
        "f": faces,
    }


def write_reference_style_pkl(model: dict, path) -> None:
    """Write ``model`` as a pickle that the reference's own ManoLayer
    constructor can open (plain ndarrays; J_regressor as scipy csc because the
    reference calls ``.todense()`` on it, MANOLayer.py:72)."""
    import scipy.sparse as sp

    dd = dict(model)
    dd["J_regressor"] = sp.csc_matrix(model["J_regressor"])
    with open(path, "wb") as fh:
        pickle.dump(dd, fh, protocol=2)


def parents_from_kintree(kintree_table: np.ndarray) -> np.ndarray:
    """Same mapping as MANOLayer.py:66-67 (id_to_col / parent), root -> -1."""
    kt = np.asarray(kintree_table)
    id_to_col = {int(kt[1, i]): i for i in range(kt.shape[1])}
    parents = np.full(kt.shape[1], -1, dtype=np.int32)
    for i in range(1, kt.shape[1]):
        parents[i] = id_to_col[int(kt[0, i])]
        if not (0 <= parents[i] < i):
            raise ValueError("kintree_table must list parents before children")
    return parents


@dataclass
class PackedMano:
    """fp32 host arrays in the exact layouts the CUDA kernels consume.

    basis      [FEAT_K, 2334]   rows: 10 shapedirs, 135 posedirs, v_template, 2 zero pad;
                                column = vertex*3 + coord           (forward GEMM B operand)
    basis_t    [2334, FEAT_K]   transpose of ``basis``                (backward GEMM B operand)
    j0         [16, 3]          J_regressor @ v_template
    jb         [16, 3, 10]      J_regressor @ shapedirs               (joint regression folded)
    pca        [nc, 45]         hands_components[:nc]
    pose_mean  [45]
    skin_w     [778, MAX_INFL]  ELL skin weights (zero padded)
    skin_b     [778, MAX_INFL]  bone id per ELL slot (int32, 0 padded)
    parents    [16] int32, depth [16] int32
    """

    nc: int
    basis: np.ndarray
    basis_t: np.ndarray
    j0: np.ndarray
    jb: np.ndarray
    pca: np.ndarray
    pose_mean: np.ndarray
    skin_w: np.ndarray
    skin_b: np.ndarray
    parents: np.ndarray
    depth: np.ndarray


def pack_mano(model: dict, nc: int) -> PackedMano:
    """Fold and lay out the constants.  All folding is done in float64 and
    rounded once to fp32 (the reference casts each raw constant to fp32 first,
    MANOLayer.py:69-75; the difference is below 1e-9 m)."""
    if not (1 <= nc <= N_POSE_AA):
        raise ValueError(f"pose_num must be in [1, {N_POSE_AA}], got {nc}")
    _check_shapes(model)
    # The reference holds every constant as fp32; fold from the fp32-rounded
    # values so that the folded regressors equal what the reference multiplies.
    vt = model["v_template"].astype(np.float32).astype(np.float64)
    sd = model["shapedirs"].astype(np.float32).astype(np.float64)
    pd = model["posedirs"].astype(np.float32).astype(np.float64)
    jr = model["J_regressor"].astype(np.float32).astype(np.float64)
    w = model["weights"].astype(np.float32)

    basis = np.zeros((FEAT_K, N_VC), dtype=np.float32)
    basis[0:N_BETAS] = sd.reshape(N_VC, N_BETAS).T
    basis[N_BETAS:N_BETAS + N_POSE_FEAT] = pd.reshape(N_VC, N_POSE_FEAT).T
    basis[FEAT_ONE] = vt.reshape(N_VC)

    j0 = (jr @ vt).astype(np.float32)
    jb = np.einsum("jv,vcs->jcs", jr, sd).astype(np.float32)

    nnz = (w != 0).sum(axis=1)
    if nnz.max() > MAX_INFL:
        raise ValueError(f"a vertex has {nnz.max()} bone influences; this build supports <= {MAX_INFL}")
    skin_w = np.zeros((N_VERTS, MAX_INFL), dtype=np.float32)
    skin_b = np.zeros((N_VERTS, MAX_INFL), dtype=np.int32)
    for v in range(N_VERTS):
        idx = np.nonzero(w[v])[0]
        skin_w[v, :idx.size] = w[v, idx]
        skin_b[v, :idx.size] = idx

    parents = parents_from_kintree(model["kintree_table"])
    depth = np.zeros(N_JOINTS, dtype=np.int32)
    for i in range(1, N_JOINTS):
        depth[i] = depth[parents[i]] + 1

    return PackedMano(
        nc=nc,
        basis=np.ascontiguousarray(basis),
        basis_t=np.ascontiguousarray(basis.T),
        j0=j0,
        jb=jb,
        pca=np.ascontiguousarray(model["hands_components"][:nc].astype(np.float32)),
        pose_mean=model["hands_mean"].astype(np.float32),
        skin_w=skin_w,
        skin_b=skin_b,
        parents=parents,
        depth=depth,
    )
