"""Flat struct-of-arrays physics model ("model blob") shared by every backend.

The reference builds its model with ``mujoco_py.load_model_from_path``
(/root/reference/hsr/mujoco_env.py:33) and keeps it inside libmujoco.  Here the
compiled model is an explicit, serialisable set of numpy arrays:

* ``Model``            – dataclass of arrays (fp64 / int32), produced by ``mjcf.compile_model``
* ``Model.to_blob()``  – little-endian byte string consumed by ``hsrb_create`` (include/hsrb.h)
* ``Model.from_blob()``– inverse, used to load the pre-compiled blobs shipped in ``blobs/``

Bodies that have no joint between them and their parent are *fused* into the parent at compile time
(rigid-body composition is exact), so ``nbody`` counts only the world plus bodies that carry joints.
Quantities MuJoCo defines on the un-fused bodies and that influence the dynamics
(``body_invweight0`` -> per-geom ``geom_invweight``; collision filtering -> static pair list)
are evaluated on the original tree before fusing.
"""
from __future__ import annotations

import dataclasses
import struct
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

MAGIC = 0x42525348  # 'HSRB'
VERSION = 3

# joint types
JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 1, 2
# geom types (same order as MuJoCo's mjtGeom for the subset in use: pair ordering depends on it)
GEOM_PLANE, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 0, 5, 6, 7
# narrowphase selector stored per pair
NP_PLANE_BOX, NP_PLANE_CONVEX, NP_BOX_BOX, NP_CONVEX_CONVEX = 0, 1, 2, 3

_F = np.float64
_I = np.int32

# (name, dtype, shape-expression evaluated against the dims dict)
_FIELDS = [
    # options
    ("opt", _F, "(16,)"),  # timestep gx gy gz impratio tolerance ls_tolerance iterations ls_iterations
    #                        mpr_tolerance mpr_iterations meaninertia (4 spare)
    ("qpos0", _F, "(nq,)"),
    # bodies (fused)
    ("body_parent", _I, "(nbody,)"),
    ("body_pos", _F, "(nbody,3)"),
    ("body_quat", _F, "(nbody,4)"),
    ("body_mass", _F, "(nbody,)"),
    ("body_ipos", _F, "(nbody,3)"),
    ("body_inertia", _F, "(nbody,6)"),  # xx yy zz xy xz yz about CoM, body frame
    ("body_jntadr", _I, "(nbody,)"),
    ("body_jntnum", _I, "(nbody,)"),
    ("body_dofadr", _I, "(nbody,)"),
    ("body_dofnum", _I, "(nbody,)"),
    # joints
    ("jnt_type", _I, "(njnt,)"),
    ("jnt_body", _I, "(njnt,)"),
    ("jnt_qposadr", _I, "(njnt,)"),
    ("jnt_dofadr", _I, "(njnt,)"),
    ("jnt_axis", _F, "(njnt,3)"),
    ("jnt_pos", _F, "(njnt,3)"),
    ("jnt_limited", _I, "(njnt,)"),
    ("jnt_range", _F, "(njnt,2)"),
    ("jnt_solref", _F, "(njnt,2)"),
    ("jnt_solimp", _F, "(njnt,5)"),
    # dofs
    ("dof_body", _I, "(nv,)"),
    ("dof_parent", _I, "(nv,)"),
    ("dof_jnt", _I, "(nv,)"),
    ("dof_damping", _F, "(nv,)"),
    ("dof_invweight0", _F, "(nv,)"),
    # actuators (<position>)
    ("act_dof", _I, "(nu,)"),
    ("act_qposadr", _I, "(nu,)"),
    ("act_gear", _F, "(nu,)"),
    ("act_kp", _F, "(nu,)"),
    ("act_ctrllimited", _I, "(nu,)"),
    ("act_ctrlrange", _F, "(nu,2)"),
    ("act_forcelimited", _I, "(nu,)"),
    ("act_forcerange", _F, "(nu,2)"),
    # colliding geoms
    ("geom_type", _I, "(ngeom,)"),
    ("geom_body", _I, "(ngeom,)"),
    ("geom_pos", _F, "(ngeom,3)"),  # geom centre in (fused) body frame
    ("geom_mat", _F, "(ngeom,9)"),  # geom orientation in body frame, row major
    ("geom_size", _F, "(ngeom,3)"),
    ("geom_rbound", _F, "(ngeom,)"),
    ("geom_aabb", _F, "(ngeom,3)"),  # half extents of the geom-frame AABB (midphase cull)
    ("geom_vertadr", _I, "(ngeom,)"),
    ("geom_vertnum", _I, "(ngeom,)"),
    ("geom_invweight", _F, "(ngeom,)"),  # body_invweight0[2*b] of the *original* body
    ("hull_vert", _F, "(nvert,3)"),  # convex hull vertices, geom frame
    # static candidate pair list
    ("pair_geom1", _I, "(npair,)"),
    ("pair_geom2", _I, "(npair,)"),
    ("pair_func", _I, "(npair,)"),
    ("pair_condim", _I, "(npair,)"),
    ("pair_friction", _F, "(npair,5)"),
    ("pair_solref", _F, "(npair,2)"),
    ("pair_solimp", _F, "(npair,5)"),
    # env-level bookkeeping
    ("block_body", _I, "(nblock,)"),  # fused body id of block{i}
    ("finger_body", _I, "(2,)"),  # fused body ids carrying hand_l/r_distal_link
    ("finger_pos", _F, "(2,3)"),  # position of those links in the fused body frame
    ("mocap_pos0", _F, "(3,)"),
]

_DIMS = ["nq", "nv", "nu", "nbody", "njnt", "ngeom", "nvert", "npair", "nblock"]


@dataclass
class Model:
    nq: int = 0
    nv: int = 0
    nu: int = 0
    nbody: int = 0
    njnt: int = 0
    ngeom: int = 0
    nvert: int = 0
    npair: int = 0
    nblock: int = 0
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)
    # host-only metadata (not serialised into the blob)
    names: Dict[str, List[str]] = field(default_factory=dict)
    info: Dict[str, object] = field(default_factory=dict)

    def __getattr__(self, item):
        arrays = object.__getattribute__(self, "arrays")
        if item in arrays:
            return arrays[item]
        raise AttributeError(item)

    # convenient named views of opt
    @property
    def timestep(self):
        return float(self.opt[0])

    @property
    def gravity(self):
        return self.opt[1:4]

    @property
    def impratio(self):
        return float(self.opt[4])

    @property
    def meaninertia(self):
        return float(self.opt[11])

    def dims(self):
        return {k: getattr(self, k) for k in _DIMS}

    def validate(self):
        dims = self.dims()
        for name, dt, shp in _FIELDS:
            want = eval(shp, {}, dims)
            a = self.arrays[name]
            assert a.dtype == dt, (name, a.dtype)
            assert a.shape == want, (name, a.shape, want)

    def to_blob(self) -> bytes:
        self.validate()
        head = struct.pack("<%di" % (2 + len(_DIMS)), MAGIC, VERSION, *[getattr(self, k) for k in _DIMS])
        head += b"\0" * (64 - len(head))
        parts = [head]
        for name, dt, _ in _FIELDS:
            raw = np.ascontiguousarray(self.arrays[name]).astype("<f8" if dt == _F else "<i4").tobytes()
            pad = (-len(raw)) % 8
            parts.append(raw + b"\0" * pad)
        return b"".join(parts)

    @classmethod
    def from_blob(cls, blob: bytes) -> "Model":
        vals = struct.unpack_from("<%di" % (2 + len(_DIMS)), blob, 0)
        if vals[0] != MAGIC or vals[1] != VERSION:
            raise ValueError("not an HSRB model blob (magic/version mismatch)")
        m = cls(**dict(zip(_DIMS, vals[2:])))
        dims = m.dims()
        off = 64
        for name, dt, shp in _FIELDS:
            shape = eval(shp, {}, dims)
            n = int(np.prod(shape))
            item = 8 if dt == _F else 4
            a = np.frombuffer(blob, dtype="<f8" if dt == _F else "<i4", count=n, offset=off).reshape(shape)
            m.arrays[name] = a.astype(dt).copy()
            off += n * item + ((-n * item) % 8)
        if off != len(blob):
            raise ValueError("model blob has trailing or missing bytes")
        return m

    def save(self, path):
        with open(path, "wb") as f:
            f.write(self.to_blob())

    def save_with_names(self, path):
        """blob + a ``.json`` sidecar with the name tables (body / joint / actuator / geom names)."""
        import json
        from pathlib import Path
        self.save(path)
        Path(str(path) + ".json").write_text(json.dumps(dict(names=self.names, mesh_inertia=self.info.get("mesh_inertia"))))

    @classmethod
    def load(cls, path) -> "Model":
        import json
        from pathlib import Path
        with open(path, "rb") as f:
            m = cls.from_blob(f.read())
        side = Path(str(path) + ".json")
        if side.exists():
            meta = json.loads(side.read_text())
            m.names = meta.get("names", {})
            m.info = {k: v for k, v in meta.items() if k != "names"}
        return m

    def replace(self, **kw):
        m = dataclasses.replace(self)
        m.arrays = dict(self.arrays)
        m.arrays.update(kw)
        return m


def field_table():
    """(name, is_float, shape-expr) list; the C side (csrc/hsrb_model.h) walks the same table order."""
    return [(n, dt == _F, s) for n, dt, s in _FIELDS]
