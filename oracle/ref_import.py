"""ORACLE helper (test infrastructure) — imports the UNMODIFIED reference: from /root/reference in the authoring
container, else from the byte-for-byte travelling copy ``oracle/_ref/`` that ``oracle/make_ref.py`` stages (git-ignored,
shipped to the GPU box by gpurun); callers must skip when ``available()`` is False.  Used to pin the numpy oracle, to
generate ``tests/golden``, as the live parity checker of the ``-m gpu`` tests and as the CPU baseline arm of
``bench.py``.  The reference modules are imported as they are, behind stub modules for its missing third-party imports
(``mano`` — mesh viewer only, ``chumpy`` — unpickling only, ``matplotlib``).
"""
from __future__ import annotations

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = "/root/reference" if os.path.isfile("/root/reference/network/sub_modules/MANOLayer.py") else _STAGED
REAL_PKL = os.path.join(REF_ROOT, "config/mano/models/MANO_RIGHT.pkl")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "network/sub_modules/MANOLayer.py"))


def real_pkl_available() -> bool:
    return os.path.isfile(REAL_PKL)


def _install_stubs():
    import numpy as np

    if "chumpy" not in sys.modules:
        ch = types.ModuleType("chumpy")
        ch_ch = types.ModuleType("chumpy.ch")
        ch_re = types.ModuleType("chumpy.reordering")

        class Ch:
            def __setstate__(self, st):
                self.__dict__.update(st)

            @property
            def r(self):
                return np.asarray(self.x)

            def __array__(self, dtype=None, copy=None):
                a = np.asarray(self.r)
                return a.astype(dtype) if dtype is not None else a

            @property
            def shape(self):
                return self.r.shape

        class Select(Ch):
            @property
            def r(self):
                return np.asarray(self.a.r).ravel()[self.idxs].reshape(self.preferred_shape)

        ch_ch.Ch = Ch
        ch.Ch = Ch
        ch_re.Select = Select
        ch.ch = ch_ch
        ch.reordering = ch_re
        sys.modules.update({"chumpy": ch, "chumpy.ch": ch_ch, "chumpy.reordering": ch_re})
    if "mano" not in sys.modules:
        m = types.ModuleType("mano")
        mu = types.ModuleType("mano.utils")
        mu.Mesh = object
        m.utils = mu
        sys.modules.update({"mano": m, "mano.utils": mu})
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mp = types.ModuleType("matplotlib")
        mpp = types.ModuleType("matplotlib.pyplot")
        mp.pyplot = mpp
        sys.modules.update({"matplotlib": mp, "matplotlib.pyplot": mpp})


def load():
    """Return a namespace with the reference's ManoLayer, ForwardKinematics,
    MPJPE, L2Loss, batch_project_xyz_to_uv and its mutable ``config`` module."""
    if not available():
        raise RuntimeError("reference sources are not present on this machine")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from network.sub_modules.MANOLayer import ManoLayer
    from network.sub_modules.forwardKinematicsLayer import ForwardKinematics
    from criterions.metrics import MPJPE
    from criterions.loss import L2Loss
    from utils.coordinate_trans import batch_project_xyz_to_uv
    from config import config

    return types.SimpleNamespace(ManoLayer=ManoLayer, ForwardKinematics=ForwardKinematics, MPJPE=MPJPE,
                                 L2Loss=L2Loss, batch_project_xyz_to_uv=batch_project_xyz_to_uv, config=config)
