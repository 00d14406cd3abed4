"""IOManager file formats: run.h5 + run.xmf written by the product's own
HDF5 writer (fv2d_b200/host/H5Lite.h, no libhdf5 in this image), checked with the independent
pure-Python reader tests/h5mini.py, which is itself pinned on a file written by the real HDF5
library (the MATLAB fixture shipped with scipy).  CPU only: the C ABI's fv2d_io_* entry points
work on host arrays.

Reference behaviour being mirrored: IOManager.h:99-398 (layout: root attributes, x / y vertex
datasets, one ite_%04d group per snapshot with rho,u,v,prs + time,iteration; XDMF footer
rewritten in place; restart from 'file.h5' or 'file.h5:/ite_NNNN')."""
import os
import struct
from pathlib import Path

import numpy as np
import pytest

import h5mini
from conftest import ROOT
from fv2d_b200 import capi

SCIPY_FIXTURE = None
try:
    import scipy.io

    _p = Path(scipy.io.__file__).parent / "matlab" / "tests" / "data" / "testhdf5_7.4_GLNX86.mat"
    if _p.exists():
        SCIPY_FIXTURE = _p
except Exception:  # pragma: no cover
    pass


def _params(tmp_path, ini="sod_x.ini", **ov):
    over = {"run.output_path": str(tmp_path), "run.output_filename": "run"}
    over.update(ov)
    return capi.params_from_ini(ROOT / "settings" / ini, over)


def _state(dev, run, k=0):
    Q = capi.init_problem(dev, run)
    rng = np.random.default_rng(100 + k)
    Q[:, dev.jbeg:dev.jend, dev.ibeg:dev.iend] += 0.01 * k + 1e-3 * rng.random((4, dev.Ny, dev.Nx))
    return Q


def _dom(dev, Q):
    return Q[:, dev.jbeg:dev.jend, dev.ibeg:dev.iend]


# ---------------------------------------------------------------------------- the reader is pinned


@pytest.mark.skipif(SCIPY_FIXTURE is None, reason="scipy's MATLAB HDF5 fixture not installed")
def test_reader_parses_a_file_written_by_the_real_hdf5_library():
    f = h5mini.File(SCIPY_FIXTURE)
    assert f.sb_off == 512 and f.base == 0x200 and (f.leaf_k, f.internal_k) == (4, 16)
    assert f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.attrs["MATLAB_class"] == "double"
    assert np.allclose(d.read().ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


@pytest.mark.skipif(SCIPY_FIXTURE is None, reason="scipy's MATLAB HDF5 fixture not installed")
def test_writer_encodes_fp64_and_dataspace_like_libhdf5(tmp_path):
    """Byte-for-byte: the IEEE binary64 datatype message and the version-1 dataspace header our
    writer emits equal the ones libhdf5 wrote into the fixture."""
    dev, run = _params(tmp_path)
    capi.io_save_solution(dev, run, _state(dev, run), 0, 0.0)
    ours, real = h5mini.File(tmp_path / "run.h5"), h5mini.File(SCIPY_FIXTURE)
    msg = lambda node, t: next(b for ty, _, b in node.messages if ty == t)
    assert msg(ours["x"], 0x0003)[:20] == msg(real["testdouble"], 0x0003)[:20]
    assert msg(ours["x"], 0x0001)[:2] + msg(ours["x"], 0x0001)[2:8] == b"\x01\x01" + bytes(6)
    assert msg(real["testdouble"], 0x0001)[0] == 1  # same dataspace message version
    # local heap and symbol-node conventions match too: "" at offset 0, names 8-byte aligned,
    # the free list terminated by H5HL_FREE_NULL (1)
    for f in (ours, real):
        size, free, addr = f.local_heap_info(f.heap)
        data = f.buf[f.base + addr:f.base + addr + size]
        assert data[:8] == bytes(8) and size % 8 == 0
        nxt, flen = struct.unpack_from("<QQ", data, free)
        assert nxt == 1 and free + flen <= size


# ---------------------------------------------------------------------------- unique-file mode


def test_unique_file_layout_and_contents(tmp_path):
    dev, run = _params(tmp_path)
    states, flag = [], False
    for it in range(5):
        Q = _state(dev, run, it)
        states.append(Q)
        flag = capi.io_save_solution(dev, run, Q, it, 0.125 * it, flag)
        assert flag is False
    f = h5mini.File(tmp_path / "run.h5")
    assert f.sb_version == 0 and f.base == 0 and f.eof == os.path.getsize(tmp_path / "run.h5")
    # name order == what HighFive::getObjectName(index) walks (IOManager.h:330-334)
    assert f.keys() == [f"ite_{k:04d}" for k in range(5)] + ["x", "y"]
    want = {"Ntx": dev.Ntx, "Nty": dev.Nty, "Nx": dev.Nx, "Ny": dev.Ny, "ibeg": dev.ibeg, "iend": dev.iend,
            "jbeg": dev.jbeg, "jend": dev.jend}
    assert list(f.attrs) == list(want) + ["problem"]  # creation order, IOManager.h:217-225
    for k, v in want.items():
        assert f.attrs[k].dtype == np.int32 and int(f.attrs[k]) == v
    assert f.attrs["problem"] == "sod_x"
    # vertex coordinates, (Ny+1) x (Nx+1), j-major (IOManager.h:227-236)
    x, y = f["x"].read(), f["y"].read()
    jj, ii = np.meshgrid(np.arange(dev.Ny + 1), np.arange(dev.Nx + 1), indexing="ij")
    assert np.array_equal(x, (ii * dev.dx + dev.xmin).ravel()) and np.array_equal(y, (jj * dev.dy + dev.ymin).ravel())
    for it, Q in enumerate(states):
        g = f[f"ite_{it:04d}"]
        assert g.keys() == ["prs", "rho", "u", "v"]
        assert list(g.attrs) == ["time", "iteration"]
        assert g.attrs["time"].dtype == np.float64 and float(g.attrs["time"]) == 0.125 * it
        assert g.attrs["iteration"].dtype == np.int32 and int(g.attrs["iteration"]) == it
        for name, fld in (("rho", 0), ("u", 1), ("v", 2), ("prs", 3)):
            d = g[name]
            assert d.shape == (dev.Nx * dev.Ny,) and d.layout[0] == "contiguous"
            assert np.array_equal(d.read().reshape(dev.Ny, dev.Nx), _dom(dev, Q)[fld])


def test_iteration_zero_truncates_an_existing_file(tmp_path):
    dev, run = _params(tmp_path)
    for it in range(3):
        capi.io_save_solution(dev, run, _state(dev, run, it), it, float(it))
    capi.io_save_solution(dev, run, _state(dev, run, 7), 0, 0.0)  # iteration == 0 -> File::Truncate
    f = h5mini.File(tmp_path / "run.h5")
    assert f.keys() == ["ite_0000", "x", "y"]


def test_many_snapshots_build_a_multi_level_btree(tmp_path):
    """> 2K_leaf * 2K_internal = 256 links force a second B-tree level; every append rewrites
    only the root group's index, never the datasets."""
    dev, run = _params(tmp_path, **{"mesh.Nx": 6, "mesh.Ny": 4})
    n, flag = 300, False
    for it in range(n):
        Q = np.full(dev.shape(), float(it))
        flag = capi.io_save_solution(dev, run, Q, it, 1e-3 * it, flag)
    f = h5mini.File(tmp_path / "run.h5")
    assert f.keys() == [f"ite_{k:04d}" for k in range(n)] + ["x", "y"]  # the reader also checks the key order
    level = f.buf[f.btree + 5]
    assert f.buf[f.btree:f.btree + 4] == b"TREE" and level == 1
    for it in (0, 1, 255, 256, 299):
        assert np.all(f[f"ite_{it:04d}/u"].read() == float(it))
        assert int(f[f"ite_{it:04d}"].attrs["iteration"]) == it


def test_xdmf_sidecar_matches_the_reference_text(tmp_path):
    """Expected text built here from the reference's format strings (IOManager.h:26-71),
    including its seek-back of sizeof(footer) = strlen + 1 (IOManager.h:274)."""
    dev, run = _params(tmp_path)
    for it in range(3):
        capi.io_save_solution(dev, run, _state(dev, run, it), it, 0.1 * it)
    head = ('<?xml version="1.0" ?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" [\n<!ENTITY file "%s:">\n<!ENTITY fdim "%d %d">\n'
            '<!ENTITY gdim "%d %d">\n<!ENTITY GridEntity \'\n<Topology TopologyType="2DSMesh" Dimensions="&gdim;"/>\n'
            '<Geometry GeometryType="X_Y">\n'
            '  <DataItem Dimensions="&gdim;" NumberType="Float" Precision="8" Format="HDF">&file;/x</DataItem>\n'
            '  <DataItem Dimensions="&gdim;" NumberType="Float" Precision="8" Format="HDF">&file;/y</DataItem>\n'
            '</Geometry>\'>\n]>\n<Xdmf Version="3.0">\n<Domain>\n'
            '  <Grid Name="TimeSeries" GridType="Collection" CollectionType="Temporal">\n    '
            ) % ("run.h5", dev.Ny, dev.Nx, dev.Ny + 1, dev.Nx + 1)
    foot = "\n  </Grid>\n</Domain>\n</Xdmf>"
    item = '<DataItem Dimensions="&fdim;" NumberType="Float" Precision="8" Format="HDF">&file;/%s%s</DataItem>'
    scalar = ('\n      <Attribute Name="%s" AttributeType="Scalar" Center="Cell">\n        ' + item
              + "\n      </Attribute>")
    vector = ('\n      <Attribute Name="%s" AttributeType="Vector" Center="Cell">\n'
              '        <DataItem Dimensions="&fdim; 2" ItemType="Function" Function="JOIN($0, $1)">\n          '
              + item + "\n          " + item + "\n        </DataItem>\n      </Attribute>")
    text = head + foot
    for it in range(3):
        name = f"ite_{it:04d}"
        grp = name + "/"
        text = text[:len(text) - (len(foot) + 1)]
        text += '\n    <Grid Name="%s" GridType="Uniform">\n      <Time Value="%f" />\n      &GridEntity;' % (name, 0.1 * it)
        text += scalar % ("rho", grp, "rho") + vector % ("velocity", grp, "u", grp, "v") + scalar % ("prs", grp, "prs")
        text += "\n    </Grid>\n    " + foot
    assert (tmp_path / "run.xmf").read_text() == text


# ---------------------------------------------------------------------------- multiple-file mode


def test_multiple_outputs_mode(tmp_path):
    dev, run = _params(tmp_path, **{"run.multiple_outputs": "true"})
    assert run.multiple_outputs == 1
    Q = _state(dev, run, 3)
    capi.io_save_solution(dev, run, Q, 3, 0.125)
    f = h5mini.File(tmp_path / "run_0003.h5")
    assert f.keys() == ["prs", "rho", "u", "v", "x", "y"]
    assert float(f.attrs["time"]) == 0.125 and int(f.attrs["iteration"]) == 3 and f.attrs["problem"] == "sod_x"
    assert np.array_equal(f["prs"].read().reshape(dev.Ny, dev.Nx), _dom(dev, Q)[3])
    xmf = (tmp_path / "run_0003.xmf").read_text()
    assert '<!ENTITY file "run_0003.h5:">' in xmf and "&file;/rho</DataItem>" in xmf and xmf.endswith("</Xdmf>")
    # restart from a per-snapshot file: time / iteration are root attributes (IOManager.h:318-324)
    run.restart_file = str(tmp_path / "run_0003.h5").encode()
    Q2, t, it, _ = capi.io_load_snapshot(dev, run)
    assert (t, it) == (0.125, 3) and np.array_equal(_dom(dev, Q2), _dom(dev, Q))


# ---------------------------------------------------------------------------- restart


def _write_series(tmp_path, n=4, ini="sod_x.ini", **ov):
    dev, run = _params(tmp_path, ini, **ov)
    states = []
    for it in range(n):
        Q = _state(dev, run, it)
        states.append(Q)
        capi.io_save_solution(dev, run, Q, it, 0.05 * it)
    return dev, run, states


def test_restart_from_last_and_from_named_iteration(tmp_path):
    dev, run, states = _write_series(tmp_path)
    other = tmp_path / "elsewhere"
    other.mkdir()
    run.output_path = str(other).encode()
    run.restart_file = str(tmp_path / "run.h5").encode()
    Q, t, it, trunc = capi.io_load_snapshot(dev, run)
    assert (t, it) == (0.05 * 3, 3) and trunc is True  # a different output file will be truncated
    assert np.array_equal(_dom(dev, Q), _dom(dev, states[3]))
    run.restart_file = (str(tmp_path / "run.h5") + ":/ite_0001").encode()
    Q, t, it, _ = capi.io_load_snapshot(dev, run)
    assert (t, it) == (0.05, 1) and np.array_equal(_dom(dev, Q), _dom(dev, states[1]))


@pytest.mark.parametrize("ini", ["sod_x.ini", "blast.ini", "kelvin_helmholtz.ini"])
def test_restart_fills_the_ghost_cells(tmp_path, ini):
    """loadSnapshot ends with fillBoundaries (IOManager.h:378-379): reflecting, periodic and
    periodic/absorbing configurations come back with the ghosts a fresh init would have."""
    dev, run = _params(tmp_path, ini, **{"mesh.Nx": 24, "mesh.Ny": 12})
    Q0 = capi.init_problem(dev, run)  # init also ends with fillBoundaries (Init.h:356-357)
    capi.io_save_solution(dev, run, Q0, 0, 0.0)
    # kelvin_helmholtz.ini ships with multiple_outputs=true: one file per snapshot
    run.restart_file = str(tmp_path / ("run_0000.h5" if run.multiple_outputs else "run.h5")).encode()
    Q, *_ = capi.io_load_snapshot(dev, run)
    assert np.array_equal(Q, Q0)


def test_restart_into_the_same_file_appends(tmp_path):
    dev, run, states = _write_series(tmp_path, n=3)
    run.restart_file = str(tmp_path / "run.h5").encode()
    Q, t, it, trunc = capi.io_load_snapshot(dev, run)
    assert it == 2 and trunc is False  # same file: keep appending (IOManager.h:300-315)
    capi.io_save_solution(dev, run, Q, it + 1, t + 0.05, trunc)
    assert h5mini.File(tmp_path / "run.h5").keys() == ["ite_0000", "ite_0001", "ite_0002", "ite_0003", "x", "y"]
    run.restart_file = (str(tmp_path / "run.h5") + ":/ite_0001").encode()
    with pytest.raises(capi.Fv2dError, match="Invalid restart_file"):
        capi.io_load_snapshot(dev, run)


def test_restart_errors(tmp_path):
    dev, run, _ = _write_series(tmp_path, n=2)
    other = tmp_path / "o"
    other.mkdir()
    dev2, run2 = _params(other, **{"mesh.Nx": 32})
    run2.restart_file = str(tmp_path / "run.h5").encode()
    with pytest.raises(capi.Fv2dError, match="different resolution"):
        capi.io_load_snapshot(dev2, run2)
    dev3, run3 = _params(other, **{"run.tend": 0.01})
    run3.restart_file = str(tmp_path / "run.h5").encode()
    with pytest.raises(capi.Fv2dError, match="greater than the end time"):
        capi.io_load_snapshot(dev3, run3)
    run3.restart_file = str(tmp_path / "missing.h5").encode()
    with pytest.raises(capi.Fv2dError):
        capi.io_load_snapshot(dev3, run3)
    junk = tmp_path / "junk.h5"
    junk.write_bytes(b"not an hdf5 file" * 100)
    run3.restart_file = str(junk).encode()
    with pytest.raises(capi.Fv2dError, match="not an HDF5 file"):
        capi.io_load_snapshot(dev3, run3)
    whole = (tmp_path / "run.h5").read_bytes()
    cut = tmp_path / "cut.h5"
    cut.write_bytes(whole[:len(whole) // 2])
    run3.restart_file = str(cut).encode()
    with pytest.raises(capi.Fv2dError, match="truncated"):
        capi.io_load_snapshot(dev3, run3)


def test_io_errors_are_reported_not_thrown(tmp_path):
    dev, run = _params(tmp_path, **{"run.output_path": str(tmp_path / "does" / "not" / "exist")})
    with pytest.raises(capi.Fv2dError) as e:
        capi.io_save_solution(dev, run, capi.init_problem(dev, run), 0, 0.0)
    assert e.value.code == 3  # FV2D_ERR_IO


# ---------------------------------------------------------------------------- the writer against libhdf5's own bytes


def _messages(path):
    """(object path, message type, flags, body) of every object-header message of a file."""
    f = h5mini.File(path)
    out = []

    def rec(node, name):
        for mtype, mflags, body in node.messages:
            out.append((name, mtype, mflags, bytes(body)))
        if node.is_group:
            for k in node.keys():
                rec(node[k], name + "/" + k)

    rec(f, "")
    return f, out


@pytest.mark.skipif(SCIPY_FIXTURE is None, reason="scipy's MATLAB HDF5 fixture not installed")
def test_writer_uses_the_encodings_the_real_hdf5_library_wrote(tmp_path):
    """No libhdf5 can open our files here (f-2 stays unpinned), but the pieces can be held against
    bytes that libhdf5 itself produced: the one file in this image written by the real library (the
    MATLAB fixture the reader is pinned on) contains a double dataset with an attribute, i.e. the same
    message kinds run.h5 is made of.  The IEEE double datatype message must be byte for byte libhdf5's,
    dataspace / attribute / symbol-table messages must use the same versions and padding rules, the
    superblock the same version, offset sizes and B-tree parameters."""
    dev, run = _params(tmp_path)
    capi.io_save_solution(dev, run, _state(dev, run), 0, 0.0)
    ours_f, ours = _messages(tmp_path / "run.h5")
    ref_f, ref = _messages(SCIPY_FIXTURE)

    assert (ours_f.sb_version, ours_f.leaf_k, ours_f.internal_k) == (ref_f.sb_version, ref_f.leaf_k, ref_f.internal_k)
    assert ours_f.free_addr == ref_f.free_addr == h5mini.UNDEF and ours_f.driver == ref_f.driver == h5mini.UNDEF

    def of_type(msgs, t):
        return [m for m in msgs if m[1] == t]

    # datatype of every float dataset: libhdf5's H5T_IEEE_F64LE encoding, all 24 bytes
    ref_double = [m[3] for m in of_type(ref, 0x0003) if m[3][0] & 0x0F == 1]
    assert len(ref_double) == 1
    ours_double = [m[3] for m in of_type(ours, 0x0003)]
    assert len(ours_double) == 6 and all(b == ref_double[0] for b in ours_double)  # x, y, rho, u, v, prs
    # dataspace: version 1, no permutation / max-dims flags beyond what libhdf5 set, 8 bytes per dimension
    for _, _, _, b in of_type(ours, 0x0001):
        rb = of_type(ref, 0x0001)[0][3]
        assert b[0] == rb[0] == 1 and b[2] == rb[2] and len(b) == 8 + 8 * b[1]
    # attributes: version 1 with name / datatype / dataspace each padded to 8 bytes, like libhdf5's
    ref_attr = of_type(ref, 0x000C)[0][3]
    assert ref_attr[0] == 1
    for _, _, _, b in of_type(ours, 0x000C):
        assert b[0] == 1 and b[1] == 0
        nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
        data = len(b) - 8 - sum((n + 7) // 8 * 8 for n in (nsz, tsz, ssz))
        assert data >= 0 and b[8 + nsz - 1] == 0  # NUL-terminated name counted in its size, as in the fixture
    rn = struct.unpack_from("<H", ref_attr, 2)[0]
    assert ref_attr[8 + rn - 1] == 0
    # groups: old-style symbol-table message (B-tree + local heap address), 16 bytes, as the fixture's root
    assert all(len(m[3]) == 16 for m in of_type(ours, 0x0011)) and len(of_type(ref, 0x0011)[0][3]) == 16
    # contiguous layout (version 3 = the HDF5 1.8 default; the 2004 fixture still has version 2) and fill
    # value message versions libhdf5 1.8 reads (1, 2)
    assert all(m[3][0] == 3 and m[3][1] == 1 for m in of_type(ours, 0x0008))
    assert all(m[3][0] in (1, 2) for m in of_type(ours, 0x0005))
    # every message body is a multiple of 8 bytes (object header version 1 alignment rule)
    assert all(len(m[3]) % 8 == 0 for m in ours) and all(len(m[3]) % 8 == 0 for m in ref)
