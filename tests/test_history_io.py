"""CPU: the history writer (nyles_b200/nylesIO.py) against the file layout of core/nylesIO.py:365-585.

The state lives on the host here (param['device'] = 'cpu'); the writer only ever sees `.tensor[idx]`.
The files are read back with scipy's netCDF reader, which checks the hand-appended record section
(record layout, numrecs patching) against an independent implementation of the format."""
import os

import numpy as np
import pytest

scipy_io = pytest.importorskip("scipy.io")


def close(f):
    """The experiment parameter "mode" is a global attribute of the file (as in the reference); scipy's reader
    lets it shadow its own `mode` field, which close() consults."""
    f.__dict__["mode"] = "r"
    f.close()


def make(tmp_path, **over):
    from nyles_b200 import grid as G, nylesIO, topology as topo, variables as V
    topo.topology = over.pop("geometry", "closed")
    procs = [1, 1, 1]
    ngbs = topo.get_neighbours([0, 0, 0], procs)
    param = dict(nx=6, ny=5, nz=4, nh=3, neighbours=ngbs, procs=procs, loc=[0, 0, 0], npx=1, npy=1, npz=1,
                 Lx=60.0, Ly=50.0, Lz=10.0, device="cpu", myrank=0,
                 datadir=str(tmp_path), expname="test_exp", mode="overwrite", timestep_history=1.0,
                 disk_space_warning=0.0, unit_length="m", unit_duration="s", n_tracers=0, simplified_grid=False,
                 include_halo=False, variables_in_history="p+p",
                 **{"a boolean variable": True, "a long list": list(range(50))})
    param.update(over)
    state = V.get_state(param)
    return param, state, G.Grid(param), nylesIO.NylesIO(param)


def test_history_file_layout_and_cadence(tmp_path):
    """The reference's own self-test (nylesIO.py:622-675): b = t at the times below, history every 1.0."""
    param, state, grid, io = make(tmp_path)
    rng = np.random.default_rng(0)
    for d in "ijk":
        state.u[d].view("i")[:] = rng.standard_normal(state.u[d].view("i").shape)
    io.init(state, grid, 0.0, 0)
    assert io.hist_path.endswith("test_exp_00_hist.nc") and os.path.isfile(io.hist_path)
    assert io.hist_variables == {"b": "b", "u": "u_i", "v": "u_j", "w": "u_k", "p": "p"}
    saved = [(0.0, 0)]
    for i, t in enumerate([0.2, 0.5, 0.7, 1.0, 1.5, 2.1, 3.0]):
        n = i + 1
        state.b.view("i")[:] = t * np.ones(state.b.view("i").shape)
        before = io.n_hist
        assert io.write(state, t, n) is False
        if io.n_hist > before:
            saved.append((t, n))
    assert saved == [(0.0, 0), (1.0, 4), (2.1, 6), (3.0, 7)]
    io.finalize(state, 3.0, 7)                      # already saved at n = 7: nothing is added
    assert io.n_hist == 4
    io.finalize(state, 3.3, 8)                      # a later final state is
    assert io.n_hist == 5

    f = scipy_io.netcdf_file(io.hist_path, "r", mmap=False)
    assert f.dimensions["t"] is None                # unlimited
    for x, n in zip("xyz", (6, 5, 4)):
        for p in ["b", "u", "v", "w", "vor_i", "vor_j", "vor_k"]:
            assert f.dimensions["%s_%s" % (x, p)] == n
    assert f.variables["b"].dimensions == ("t", "z_b", "y_b", "x_b")
    assert f.variables["u"].dimensions == ("t", "z_u", "y_u", "x_u")
    assert f.variables["w"].dimensions == ("t", "z_w", "y_w", "x_w")
    assert f.variables["b"].units == b"m s-2" and f.variables["u"].units == b"m2 s-1" and f.variables["t"].units == b"s"
    assert f.variables["b"].long_name == b"buoyancy" and f.variables["v"].long_name == b"covariant velocity y-component"
    assert np.array_equal(f.variables["t"][:], [0.0, 1.0, 2.1, 3.0, 3.3])
    assert np.array_equal(f.variables["n"][:], [0, 4, 6, 7, 8])
    assert f.variables["b"].shape == (5, 4, 5, 6)
    assert np.all(f.variables["b"][0] == 0.0) and np.all(f.variables["b"][1] == 1.0) and np.all(f.variables["b"][2] == 2.1)
    assert np.array_equal(f.variables["u"][3], state.u["i"].tensor.numpy())
    assert np.array_equal(f.variables["w"][4], state.u["k"].tensor.numpy())
    # coordinates: cell centres and the staggered points (grid.py:51-135)
    assert np.allclose(f.variables["x_b"][:], (np.arange(6) + 0.5) * 10.0)
    assert np.allclose(f.variables["x_u"][:], (np.arange(6) + 1.0) * 10.0)
    assert np.allclose(f.variables["z_w"][:], (np.arange(4) + 1.0) * 2.5)
    assert np.allclose(f.variables["y_vor_i"][:], (np.arange(5) + 1.0) * 10.0)
    # experiment parameters as global attributes (nylesIO.py:136-160)
    assert f.Lx == 60.0 and f.nx == 6 and f.expname == b"test_exp"
    assert f.a_boolean_variable == b"True"
    long_list = f.a_long_list.decode()
    assert long_list.startswith("<class 'list'>: [0, 1, 2") and long_list.endswith("...") and len(long_list) == 100
    close(f)


def test_simplified_grid_halo_and_variable_selection(tmp_path):
    param, state, grid, io = make(tmp_path, geometry="perio_xyz", simplified_grid=True, include_halo=True,
                                  variables_in_history=["b", "vor", "U"], expname="perio", mode="count")
    state.vor["j"].view("i")[:] = 2.0
    io.init(state, grid, 0.0, 0)
    assert io.output_directory.endswith("perio_0000")
    f = scipy_io.netcdf_file(io.hist_path, "r", mmap=False)
    assert (f.dimensions["x"], f.dimensions["y"], f.dimensions["z"]) == (12, 11, 10)      # halos included
    assert set(f.variables) == {"n", "t", "x", "y", "z", "b", "vor_i", "vor_j", "vor_k", "U", "V", "W"}
    assert f.variables["vor_j"].dimensions == ("t", "z", "y", "x") and np.all(f.variables["vor_j"][0] == 2.0)
    assert np.allclose(f.variables["x"][:], (np.arange(12) + 0.5 - 3) * 10.0)
    close(f)
    # a second experiment with the same name gets the next number (nylesIO.py:170-178)
    _, _, _, io2 = make(tmp_path, geometry="perio_xyz", expname="perio", mode="count")
    assert io2.output_directory.endswith("perio_0001")
    io.save_array_3D(np.arange(24.0).reshape(2, 3, 4), "mask", "a test array")
    g = scipy_io.netcdf_file(os.path.join(io.output_directory, "perio_0000_00_mask.nc"), "r", mmap=False)
    assert np.array_equal(g.variables["mask"][:], np.arange(24.0).reshape(2, 3, 4)) and g.variables["mask"].long_name == b"a test array"
    close(g)


def test_bad_selection_and_disabled_output(tmp_path):
    param, state, grid, io = make(tmp_path, variables_in_history=["b", "nonsense"])
    with pytest.raises(ValueError):
        io.init(state, grid, 0.0, 0)
    param, state, grid, io = make(tmp_path, variables_in_history=[])
    with pytest.raises(ValueError):
        io.init(state, grid, 0.0, 0)
    param, state, grid, io = make(tmp_path, datadir="")
    io.init(state, grid, 0.0, 0)
    assert io.write(state, 5.0, 1) is False and io.n_hist == 0 and io.hist_path is None
    io.finalize(state, 5.0, 1)


def test_join_z_slab_files(tmp_path):
    """Two z slabs written as two ranks would write them, joined into one global file (tools/join.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import join as J
    from nyles_b200 import grid as G, nylesIO, topology as topo, variables as V
    topo.topology = "closed"
    procs = [2, 1, 1]
    rng = np.random.default_rng(4)
    full = rng.standard_normal((3, 8, 5, 6))                 # (snapshot, z, y, x) of the whole domain
    for rank in range(2):
        loc = topo.rank2loc(rank, procs)
        ngbs = topo.get_neighbours(loc, procs)
        param = dict(nx=6, ny=5, nz=4, global_nx=6, global_ny=5, global_nz=8, nh=3, neighbours=ngbs, procs=procs, loc=loc,
                     npx=1, npy=1, npz=2, Lx=6.0, Ly=5.0, Lz=8.0, device="cpu", myrank=rank, geometry="closed",
                     datadir=str(tmp_path), expname="slabs", mode="overwrite", timestep_history=1.0,
                     disk_space_warning=0.0, unit_length="m", unit_duration="s", n_tracers=0, simplified_grid=False,
                     include_halo=True, variables_in_history=["b", "u"])
        state, grid, io = V.get_state(param), G.Grid(param), nylesIO.NylesIO(param)
        k0, k1 = state.b.domainindices[:2]
        for n in range(3):
            state.u["k"].view("i")[k0:k1] = full[n, rank * 4:(rank + 1) * 4]
            if n == 0:
                io.init(state, grid, 0.0, 0)
            else:
                io.write(state, float(n), n)
        io.finalize(state, 2.0, 2)
    out = J.join(os.path.join(str(tmp_path), "slabs"), "w")
    f = scipy_io.netcdf_file(out, "r", mmap=False)
    assert f.variables["w"].shape == (3, 8, 5, 6)
    assert np.array_equal(f.variables["w"][:], full.astype(np.float32))
    assert np.allclose(f.variables["z"][:], np.arange(8) + 1.0) and np.allclose(f.variables["x"][:], np.arange(6) + 0.5)
    assert np.array_equal(f.variables["t"][:], [0.0, 1.0, 2.0])
    f.close()
    p = J.read_param(os.path.join(str(tmp_path), "slabs", "slabs_00_hist.nc"))
    assert p["procs"] == [2, 1, 1] and p["include_halo"] is True and p["variables_in_history"] == ["b", "u"]
