"""Reader / writer of the reference's ``.bin`` primitive files.

``Trajectories::save_file_boost / load_file_boost`` (dynobench/src/motions.cpp, declared at
dynobench/include/dynobench/motions.hpp:289-290) stream the object through a
``boost::archive::binary_oarchive``.  What ends up in the file, for the types involved
(motions.hpp:26-60 ``Eigen::VectorXd`` as ``std::vector<double>``, :257-266
``Trajectory::serialize``, :283-287 ``Trajectories::serialize``, :463
``BOOST_CLASS_VERSION(Trajectory, 1)``):

  header   u64 22 | "serialization::archive" | u16 library version | sizeof(int), sizeof(long),
           sizeof(float), sizeof(double) as 4 bytes | u32 1 (endianness probe)
  object   the FIRST time a class is met: u8 tracking flag + u32 class version; never again
  vector   u64 count | u32 item_version (class types only) | elements
  vector<double>  u64 count | the doubles, raw (no class record: not a tracked type)
  Trajectory      states, actions, cost f64, feasible u8, start, goal, info (u64 length + bytes,
                  class version > 0 only)

Boost itself is not needed (and not installed here); the layout is pinned by the
``.bin <-> .yaml`` and ``.bin <-> .msgpack`` file pairs shipped in the reference tree
(tests/golden/formats, tests/test_formats.py).
"""
from __future__ import annotations

import struct
from typing import List, Tuple

import numpy as np

_MAGIC = b"serialization::archive"
_TRAJ_VERSION = 1  # BOOST_CLASS_VERSION(dynobench::Trajectory, 1)


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.o = data, 0
        self.seen = set()

    def take(self, n: int) -> bytes:
        if self.o + n > len(self.d):
            raise RuntimeError("boost archive: unexpected end of file")
        b = self.d[self.o:self.o + n]
        self.o += n
        return b

    def u8(self) -> int:
        return self.take(1)[0]

    def u16(self) -> int:
        return struct.unpack("<H", self.take(2))[0]

    def u32(self) -> int:
        return struct.unpack("<I", self.take(4))[0]

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def f64(self) -> float:
        return struct.unpack("<d", self.take(8))[0]

    def class_info(self, name: str) -> int:
        """tracking flag + class version, present only the first time a class is met."""
        if name in self.seen:
            return -1
        self.seen.add(name)
        self.u8()
        return self.u32()

    def vector_double(self) -> np.ndarray:
        n = self.u64()
        return np.frombuffer(self.take(8 * n), dtype="<f8").copy()

    def eigen_vector(self) -> np.ndarray:
        self.class_info("Eigen::VectorXd")
        return self.vector_double()

    def vector_of_eigen(self) -> np.ndarray:
        self.class_info("std::vector<Eigen::VectorXd>")
        n = self.u64()
        self.u32()  # item_version
        rows = [self.eigen_vector() for _ in range(n)]
        return np.array(rows) if rows else np.zeros((0, 0))


def read_trajectories_boost(path: str) -> List[Tuple[np.ndarray, np.ndarray, float]]:
    """[(states [T, nx], actions [T-1, nu], cost)] of a ``Trajectories`` archive."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    if r.u64() != len(_MAGIC) or r.take(len(_MAGIC)) != _MAGIC:
        raise RuntimeError("not a boost binary archive: " + path)
    r.u16()  # library version
    if r.take(4) != bytes([4, 8, 4, 8]):
        raise RuntimeError("boost archive written on an incompatible platform: " + path)
    if r.u32() != 1:
        raise RuntimeError("boost archive with foreign endianness: " + path)
    r.class_info("Trajectories")
    r.class_info("std::vector<Trajectory>")
    n = r.u64()
    r.u32()  # item_version
    traj_version = _TRAJ_VERSION
    out = []
    for _ in range(n):
        v = r.class_info("Trajectory")
        if v >= 0:
            traj_version = v
        states = r.vector_of_eigen()
        actions = r.vector_of_eigen()
        cost = r.f64()
        r.u8()  # feasible
        r.eigen_vector()  # start
        r.eigen_vector()  # goal
        if traj_version > 0:
            r.take(r.u64())  # info
        out.append((states, actions, cost))
    return out


class _Writer:
    def __init__(self):
        self.parts, self.seen = [], set()

    def raw(self, b: bytes):
        self.parts.append(b)

    def class_info(self, name: str, version: int = 0):
        if name not in self.seen:
            self.seen.add(name)
            self.raw(struct.pack("<BI", 0, version))

    def eigen_vector(self, v):
        self.class_info("Eigen::VectorXd")
        v = np.ascontiguousarray(v, dtype="<f8").reshape(-1)
        self.raw(struct.pack("<Q", v.shape[0]) + v.tobytes())

    def vector_of_eigen(self, m):
        self.class_info("std::vector<Eigen::VectorXd>")
        self.raw(struct.pack("<QI", len(m), 0))
        for row in m:
            self.eigen_vector(row)


def write_trajectories_boost(path: str, trajs) -> None:
    """Inverse of :func:`read_trajectories_boost`; ``trajs`` = [(states, actions, cost)]."""
    w = _Writer()
    w.raw(struct.pack("<Q", len(_MAGIC)) + _MAGIC + struct.pack("<H", 17) + bytes([4, 8, 4, 8]) + struct.pack("<I", 1))
    w.class_info("Trajectories")
    w.class_info("std::vector<Trajectory>")
    w.raw(struct.pack("<QI", len(trajs), _TRAJ_VERSION))
    for states, actions, cost in trajs:
        states, actions = np.asarray(states, dtype=np.float64), np.asarray(actions, dtype=np.float64)
        w.class_info("Trajectory", _TRAJ_VERSION)
        w.vector_of_eigen(states)
        w.vector_of_eigen(actions)
        w.raw(struct.pack("<dB", float(cost), 1))
        w.eigen_vector(states[0])
        w.eigen_vector(states[-1])
        w.raw(struct.pack("<Q", 0))  # info: empty string
    with open(path, "wb") as f:
        f.write(b"".join(w.parts))
