"""CPU test of the library's host-side block geometry (parallelpoissonsolver_b200/csrc/geometry.hpp, pure host C++) against
the oracle, which is pinned bit for bit to the unmodified reference's BlockGrid (blockGrid.hpp:15-366): rank -> location,
local sizes, index limits, boundary / communication flags, eigenvalues, for DIM = 1, 2 and 3."""
import os
import shutil
import subprocess

import pytest

from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    cxx = shutil.which("g++")
    if cxx is None or not os.path.isdir(CUDA_INC):
        pytest.skip("needs g++ and the CUDA headers")
    out = str(tmp_path_factory.mktemp("geom") / "geometry_check")
    subprocess.run([cxx, "-std=c++17", "-O1", "-I" + CUDA_INC, os.path.join(HERE, "cpp", "geometry_check.cpp"), "-o", out], check=True)
    return out


CASES = [
    (3, (128, 128, 256), (1, 1, 1), (0, 1, 0, 1, 0, 1)),
    (3, (128, 128, 256), (4, 4, 4), (0, 1, 0, 1, 0, 1)),
    (3, (24, 20, 28), (3, 2, 1), (1, 0, 0, 1, 1, 1)),
    (3, (67, 9, 12), (1, 1, 4), (0, 0, 0, 0, 0, 0)),
    (3, (25, 21, 29), (2, 2, 2), (0, 1, 0, 1, 0, 1)),    # the rank grid does not divide the grid: nlocal = npglobal / nranks truncates (blockGrid.hpp:165)
    (3, (6, 6, 6), (2, 2, 2), (0, 0, 0, 0, 0, 0)),       # smallest blocks: 3 points per axis
    (2, (24, 20, 1), (3, 2, 1), (0, 1, 0, 1, 0, 1)),
    (2, (40, 36, 1), (2, 3, 1), (0, 0, 0, 0, 0, 0)),
    (2, (40, 36, 7), (1, 1, 1), (1, 1, 1, 0, 0, 0)),     # npglobal[2] is ignored for DIM = 2 (blockGrid.hpp:166-167)
    (1, (48, 1, 1), (4, 1, 1), (0, 1, 0, 1, 0, 1)),
    (1, (48, 5, 9), (1, 1, 1), (1, 0, 0, 0, 0, 0)),
]


@pytest.mark.parametrize("dim,np_,nranks,bcs", CASES)
def test_block_geometry_matches_the_oracle(exe, dim, np_, nranks, bcs):
    ds = (0.1, 0.12, 0.09)
    o = po.Oracle(po.make_config(np_, nranks, ds=ds, bcs=bcs, dim=dim))
    for rank in range(o.world):
        r = subprocess.run([exe, str(dim), *map(str, np_), *map(str, nranks), *map(str, bcs), *map(repr, ds), str(rank)],
                           check=True, capture_output=True, text=True)
        got = {l.split()[0]: l.split()[1:] for l in r.stdout.splitlines()}
        bi = o.block(rank)
        assert list(map(int, got["loc"])) == list(bi.loc)
        assert list(map(int, got["nlocal"])) == list(bi.nlocal)
        assert list(map(int, got["nguards"])) == list(bi.nguards)
        assert list(map(int, got["limits_data"])) == list(bi.limits_data)
        assert list(map(int, got["limits_solver"])) == list(bi.limits_solver)
        assert list(map(int, got["has_boundary"])) == list(bi.has_boundary)
        assert list(map(int, got["has_comm"])) == list(bi.has_comm)
        assert int(got["ntot"][0]) == bi.ntot
        eg, el = o.eigenvalues(rank)
        assert [float(v) for v in got["eig"]] == [eg[0], eg[1], el[0], el[1]]      # bit-identical (same expression, same libm)
        # device layout: rows are padded to 128 B and the first data cell of a row is 128-byte aligned
        pitch, first = int(got["pitch"][0]), int(got["pitch"][6])
        assert pitch % 16 == 0 and (first % 16) == 0 and pitch >= bi.nlocal[0] + 2 + 15
    o.close()
