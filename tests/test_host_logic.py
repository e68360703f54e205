"""CPU tests: C-ABI surface, parameters, topology, storage/views and grid of the host layer."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nyles_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ny_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/nyles_b200.h declares."""
    from nyles_b200 import lib
    names = header_symbols()
    assert len(names) >= 30
    L = lib.load()
    for n in names:
        assert isinstance(getattr(L, n), ctypes._CFuncPtr), n
    assert sorted(lib.EXPORTED) == names, "lib.py prototypes and the header disagree"
    assert L.ny_version() >= 100


def test_no_device_means_loud_failure():
    import torch
    from nyles_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(lib.NylesB200Error):
        lib.context()
    h = ctypes.c_void_p()
    assert lib.load().ny_init(0, ctypes.byref(h)) < 0 and lib.load().ny_last_error()


def test_parameters_defaults_and_validation():
    """The thirteen failure modes exercised by core/parameters.py:344-427, on the rewritten class."""
    from nyles_b200.parameters import UserParameters, UserParameterError, InextensibleDict
    InextensibleDict.unfreeze()
    p = UserParameters()
    p.check()
    v = p.view_parameters()
    assert v["modelname"] == "LES" and v["geometry"] == "closed" and v["nh"] == 3
    assert v["timestepping"] == "LFAM3" and v["global_nx"] == 64 and v["cfl"] == 1.0 and v["orderVF"] == 5
    assert p.help("cfl") and p.possible_values("geometry")[0] == "closed"
    with pytest.raises(ValueError):
        p.help("nonsense")
    with pytest.raises(UserParameterError):
        p.model["newkey"] = 1

    def bad(cat, key, value):
        InextensibleDict.unfreeze()
        q = UserParameters()
        getattr(q, cat)[key] = value
        with pytest.raises(UserParameterError):
            q.check()

    bad("model", "geometry", "open")
    bad("model", "Lx", 0.0)
    bad("model", "Lx", "1")
    bad("model", "n_tracers", -1)
    bad("time", "timestepping", "RK4")
    bad("time", "dt", -0.1)
    bad("discretization", "global_nx", 100)
    bad("discretization", "orderA", 7)
    bad("MPI", "npx", 3)
    bad("MPI", "npz", 128)          # more sub-domains than grid points
    bad("IO", "expname", "a/b")
    bad("IO", "variables_in_history", "everything")
    bad("animation", "iterations_per_frame", 0)
    InextensibleDict.unfreeze()
    q = UserParameters()
    q.discretization["global_nx"] = 96        # 3 * 2^n is accepted by the parameter check
    q.check()
    q.freeze()
    with pytest.raises(UserParameterError):
        q.time["cfl"] = 0.5
    InextensibleDict.unfreeze()


def test_topology_neighbours_and_extents():
    from nyles_b200 import topology as topo
    procs = [2, 1, 1]
    topo.topology = "closed"
    n0 = topo.get_neighbours([0, 0, 0], procs)
    n1 = topo.get_neighbours([1, 0, 0], procs)
    assert n0 == {(1, 0, 0): 1} and n1 == {(-1, 0, 0): 0}
    size, dom = topo.get_variable_shape([8, 8, 8], n0, 3)
    assert size == [11, 8, 8] and dom == (0, 8, 0, 8, 0, 8)
    size, dom = topo.get_variable_shape([8, 8, 8], n1, 3)
    assert size == [11, 8, 8] and dom == (3, 11, 0, 8, 0, 8)
    topo.topology = "perio_xyz"
    n = topo.get_neighbours(0, [1, 1, 1])
    assert len(n) == 26 and set(n.values()) == {0}
    # symmetric connectivity (core/mpi/topology.py:224-250) on a 4-slab periodic column
    procs = [4, 1, 1]
    for r in range(4):
        for d, other in topo.get_neighbours(r, procs).items():
            back = topo.get_neighbours(other, procs)
            assert back[tuple(-c for c in d)] == r
    assert topo.rank2loc(3, [4, 1, 1]) == [3, 0, 0] and topo.loc2rank([3, 0, 0], [4, 1, 1]) == 3


def test_scalar_views_alias_one_buffer():
    from nyles_b200 import variables as V, topology as topo
    topo.topology = "perio_xy"
    p = dict(nx=4, ny=5, nz=6, nh=3, neighbours=topo.get_neighbours(0, [1, 1, 1]), device="cpu")
    s = V.Scalar(p, "buoyancy", "b", "")
    assert tuple(s.tensor.shape) == (6, 11, 10) and s.domainindices == (0, 6, 3, 8, 3, 7)
    assert s.view("i").shape == (6, 11, 10) and s.view("j").shape == (10, 6, 11) and s.view("k").shape == (11, 10, 6)
    assert s.flipview("i").shape == s.view("j").shape and s.flipview("k").shape == s.view("i").shape
    vi = s.view("i")
    vi[:] = np.arange(6 * 11 * 10, dtype=float).reshape(6, 11, 10)
    # same element through the three index orders, no copy involved
    assert s.view("j")[7, 2, 4] == s.view("i")[2, 4, 7] == s.view("k")[4, 7, 2]
    vk = s.view("k")
    vk[4, 7, 2] = -1.0
    assert s.tensor[2, 4, 7].item() == -1.0
    vi *= 2.0
    assert s.tensor[2, 4, 7].item() == -2.0
    assert isinstance(np.asarray(vi), np.ndarray) and (vi + 1.0).shape == (6, 11, 10)
    st = V.get_state(p)
    assert st.get_prognostic_scalars() == ["b", "u_i", "u_j", "u_k"]
    assert st.get("u_j") is st.u["j"]
    ds = st.duplicate_prognostic_variables()
    assert sorted(ds.toc) == ["b", "u"]
    with pytest.raises(ValueError):
        s.view("x")


def test_grid_matches_reference_formulas():
    from nyles_b200 import grid as G, topology as topo
    topo.topology = "perio_xyz"
    p = dict(nx=8, ny=4, nz=4, npx=1, npy=1, npz=1, Lx=2.0, Ly=1.0, Lz=1.0, nh=3,
             neighbours=topo.get_neighbours(0, [1, 1, 1]), loc=[0, 0, 0])
    g = G.Grid(p)
    assert g.dx == 0.25 and g.idx2 == 16.0 and g.ids2["k"] == 16.0
    x = g.x_b.view("i")
    assert x.shape == (10, 10, 14)
    assert np.allclose(x[0, 0, :], (np.arange(14) + 0.5 - 3) * 0.25)
    assert np.allclose(g.x_vel["i"].view("i")[0, 0, :], x[0, 0, :] + 0.125)
    assert np.allclose(g.z_vor["i"].view("i")[:, 0, 0], (np.arange(10) + 0.5 - 3) * 0.25 + 0.125)
    assert g.y_b.view("j").shape == (14, 10, 10)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/nyles_b200.h must compile as C99 and as C++ on its own (no torch, no CUDA
    types in the signatures)."""
    import subprocess
    import tempfile
    hdr = os.path.join(ROOT, "include", "nyles_b200.h")
    for compiler, std, ext in (("gcc", "-std=c99", ".c"), ("g++", "-std=c++11", ".cpp")):
        with tempfile.NamedTemporaryFile("w", suffix=ext, delete=False) as f:
            f.write('#include "%s"\nint main(void) { return 0; }\n' % hdr)
        env = dict(os.environ)
        env.pop("CC", None)
        out = subprocess.run([compiler, std, "-Wall", "-Werror", "-fsyntax-only", f.name], capture_output=True, text=True, env=env)
        os.unlink(f.name)
        assert out.returncode == 0, out.stderr


def build_c_example(out):
    import subprocess
    env = dict(os.environ)
    env.pop("CC", None)
    libdir = os.path.join(ROOT, "nyles_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L", libdir, "-lnyles_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", out]
    return subprocess.run(cmd, capture_output=True, text=True, env=env)


def test_c_example_links_against_the_library(tmp_path):
    """A host that is neither Python nor torch: examples/c_abi_example.c builds against include/nyles_b200.h and
    links with libnyles_b200.so (it is RUN by the GPU suite)."""
    from nyles_b200 import lib
    lib.load()
    r = build_c_example(str(tmp_path / "c_abi_example"))
    assert r.returncode == 0, r.stderr


def test_sub_views_are_bound_to_the_scalar_not_to_a_buffer():
    """b.view('i')[k0:k1] keeps a chain of indices and re-applies it to the Scalar's current tensor: the fused step
    rotates buffers, and views held by experiment scripts must keep showing the field (CPU storage, no kernels)."""
    import torch
    from nyles_b200 import variables as var
    param = dict(nx=8, ny=6, nz=4, nh=3, neighbours={}, device="cpu")
    s = var.Scalar(param, "buoyancy", "b", "L.T^-2", True)
    s.tensor[:] = torch.arange(s.tensor.numel(), dtype=torch.float64).reshape(s.tensor.shape)
    sub = s.view("i")[1:3][:, 2:5]
    subj = s.view("j")[2:6]
    first = np.asarray(sub).copy()
    assert np.array_equal(first, s.tensor.numpy()[1:3][:, 2:5])
    s.tensor = s.tensor + 1000.0                          # what LES.rhs_step does: another buffer becomes the field
    assert np.array_equal(np.asarray(sub), first + 1000.0)
    assert np.array_equal(np.asarray(subj), s.tensor.permute(2, 0, 1).numpy()[2:6])
    sub[:] = -1.0                                          # writes land in the live field
    assert float(s.tensor[1:3, 2:5].max()) == -1.0 and float(s.tensor[0].min()) >= 1000.0
    assert isinstance(s.view("i")[0, 0, 0], float)


def test_bench_workloads_and_cpu_samples():
    """bench.py's workload table: the five BASELINE configs map to grids / models as SURVEY.md 8(d) lists them, the
    CPU sample of a workload is the same set-up on a smaller box, and the forcing of the plume offers both protocols."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w = bench.workload("weak512", 8)
    assert (w["nx"], w["ny"], w["nz"]) == (1024, 1024, 1024) and w["geometry"] == "closed" and w["modelname"] == "LES"
    w = bench.workload("tgv256", 1)
    assert w["geometry"] == "perio_xyz" and w["modelname"] == "Euler3d" and abs(w["dt_max"] - 0.05) < 1e-15
    w = bench.workload("plume", 8)
    assert (w["nx"], w["ny"], w["nz"]) == (1024, 1024, 512) and w["rotating"] and w["forced"] and w["scaling"] == "strong"
    assert abs(w["nx"] * w["dx"] - 16.0) < 1e-12 and abs(w["nz"] * w["dx"] - 8.0) < 1e-12
    w = bench.workload("rt512strong", 2)
    assert (w["nx"], w["ny"], w["nz"]) == (512, 512, 512) and w["scaling"] == "strong"
    assert bench.workload("lock", 1)["nx"] == 128
    for wl, sample in (("weak512", "rayleigh-taylor 256^3"), ("tgv256", "taylor-green vortex 128^3"),
                       ("plume", "turbulent plume 256x256x128"), ("lock", "lock-exchange")):
        a = argparse.Namespace(cpu_sample="auto", workload=wl, gpus=1)
        s, same = bench.cpu_sample(a)
        assert s["name"].startswith(sample), (wl, s["name"])
        assert same == (wl == "lock")
    s, same = bench.cpu_sample(argparse.Namespace(cpu_sample="same", workload="weak512", gpus=1))
    assert same and s["nx"] == 512
    w = bench.workload("plume256", 1)
    x = (np.arange(w["nx"]) + 0.5) * w["dx"]
    z = (np.arange(w["nz"]) + 0.5) * w["dx"]
    f = bench.PlumeForcing(w, x, x, z)
    assert f.Q.shape == (w["nz"], w["ny"], w["nx"]) and f.Q.max() <= 0.1 and f.Q.min() >= 0.0
    assert f.device_tendencies(None, 0.0)["b"] is f.Q
    b, u, v = bench.initial_condition(w, x, x, z, 0)
    assert u is None and abs(b[0, 0, 0] - 0.1 * (z[0] / 8.0 - 0.5)) < 1e-15
    assert bench.bytes_per_cell_step(w, 8.0) == 1080.0 + 199.0 * 8.0


def test_ncu_facts_are_stamped_with_the_kernel_sources():
    """profiles/ncu_traffic.json carries a hash of nyles_b200/csrc; bench.py only quotes DRAM traffic and fp64
    instruction counts taken from the sources the library was built from."""
    import importlib.util
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from ncu_traffic import source_stamp
    stamp = source_stamp()
    assert len(stamp) == 16 and stamp == source_stamp()
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    d = json.load(open(path))
    assert {"stamp", "families", "fp64"} <= set(d)
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    facts = bench.ncu_facts()
    assert facts["stale"] == (facts.get("stamp") != stamp)


def test_committed_ncu_facts_were_captured_from_these_kernel_sources():
    """The roofline line quotes profiles/ncu_traffic.json only when its stamp is the hash of nyles_b200/csrc as it
    stands: a kernel edit without a fresh ncu capture (tools/gpu_round.sh) turns this test red before it ships."""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from ncu_traffic import source_stamp
    d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert d["stamp"] == source_stamp()
    assert d["families"]["rhs_momentum"] > 0 and d["fp64"]["rhs_momentum"]["dp_instr_per_cell"] > 0


def test_selfcheck_cases_cover_the_three_topologies():
    from nyles_b200 import selfcheck
    for world in (2, 4, 8):
        cs = selfcheck.cases(world)
        assert [c["geometry"] for c in cs] == ["closed", "perio_xyz", "perio_xy"]
        assert all(c["n"] == (64, 64, 64 * world) for c in cs)
        assert cs[1]["modelname"] == "Euler3d" and cs[2].get("rotating")
