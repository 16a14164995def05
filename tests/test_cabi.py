"""The drop-in boundary: the shared library loads and exports every symbol include/vecode_b200.h declares; the host
mirror binds all of them; nothing in the product path imports the oracle. No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vecode_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(vo):
    if not os.path.exists(vo.SO_PATH):
        vo.build()
    lib = ctypes.CDLL(vo.SO_PATH)
    names = _declared()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_binding_covers_header(vo):
    assert sorted(vo._cabi.SIGNATURES) == _declared()
    assert vo._cabi.lib().vo_version() == 100


def test_no_cpu_fallback(vo):
    """Without a CUDA device a context cannot be created and the error says so."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vo.VecOdeError) as e:
        vo.Context(0)
    assert "no CPU fallback" in e.value.msg


def test_tableau_handles_are_host_objects(vo, oracle):
    """Tableaux need no device: from_slices round-trips, and the built-ins equal the oracle's constants bit for bit."""
    for name, idx in oracle.TABLEAU_ID.items():
        ac, b, be, s = vo.ButcherTableu.builtin(name).arrays()
        oac, ob, obe, os_ = oracle.builtin_tableau(idx)
        assert s == os_ and list(ac) == list(oac) and list(b) == list(ob)
        assert (be is None) == (obe is None) and (be is None or list(be) == list(obe))
    t = vo.ButcherTableu.from_slices([0, 0, 0.5, 0.5], [0.0, 1.0], None, 2)
    assert t.num_stages() == 2 and t.arrays()[2] is None
    with pytest.raises(vo.VecOdeError):
        vo.ButcherTableu.from_slices([0.0] * 3, [1.0, 0.0], None, 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vec-ode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                for line in open(os.path.join(dirpath, f)):
                    code = line.split("//")[0].split("#", 1)[0] if not line.lstrip().startswith("#include") else line
                    bad = ("oracle" in code) and any(k in code for k in ("import", "#include", "CDLL", "dlopen", "subprocess"))
                    assert not bad and "libvecode_oracle" not in code, (f, line)


def _build_harness(tmp_path):
    """g++ on tests/harness/vo_harness.cpp against the shared object: what a compiled host does, no Python on the data path."""
    import subprocess
    exe = str(tmp_path / "vo_harness")
    so_dir = os.path.join(ROOT, "vec-ode_b200")
    cmd = ["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "harness", "vo_harness.cpp"), "-o", exe,
           "-L", so_dir, "-lvecode_b200", f"-Wl,-rpath,{so_dir}", "-Wl,-rpath-link,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_harness_links_and_fails_loudly_without_a_gpu(tmp_path):
    """The C++ harness compiles against include/vecode_b200.h as C++ and links every entry point it uses. Without a CUDA device the
    library must refuse (exit code 3 of the harness, message 'no CPU fallback'); with one, the harness runs to 'harness ok'."""
    import subprocess
    import torch
    exe = _build_harness(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "harness ok" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_harness_on_the_gpu(tmp_path):
    """The reference's own three tests (src/impls/nalgebra.rs:52-107) and a bit-exact Lorenz-63 RK4 sweep, driven from C++."""
    import subprocess
    exe = _build_harness(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and "harness ok" in r.stdout and "bit-exact vs restatement: yes" in r.stdout, r.stdout + r.stderr
    assert "gathered ensemble == single-ctx solve: yes" in r.stdout  # vo_group_* (NCCL inside the library) from plain C++
    import torch
    if torch.cuda.device_count() >= 2:  # one process driving two GPUs
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, VO_HARNESS_GPUS="2"))
        print(r.stdout)
        assert r.returncode == 0 and "group of 2 GPU(s)" in r.stdout and "gathered ensemble == single-ctx solve: yes" in r.stdout, r.stdout + r.stderr


def test_generated_rust_ffi_is_current_and_complete():
    """bindings/rust/vecode_b200_sys.rs (tools/gen_rust_ffi.py): the raw `extern "C"` half of the Rust binding a maintainer of the
    reference would add. No rustc here, so the check is textual: regenerating gives the committed file, and every function the header
    declares has exactly one `pub fn`."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    committed = open(mod.OUT).read()
    text, names = mod.main()
    assert text == committed, "bindings/rust/vecode_b200_sys.rs is stale: run python tools/gen_rust_ffi.py"
    assert sorted(names) == _declared()
    for n in names:
        assert len(re.findall(rf"pub fn {n}\(", text)) == 1
    assert "pub const VO_ERR_BAD_ARG: i32 = -1;" in text and "pub struct VoGroupStats" in text


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/vecode_b200.h must compile as C99 (no C++-isms, no torch or CUDA types in the signatures) and a C
    program must link against the shared object."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "vecode_b200.h"\nint main(void) { vo_step_result r; vo_group_stats g; (void)r; (void)g; return vo_version() > 0 ? 0 : 1; }\n')
    so_dir = os.path.join(ROOT, "vec-ode_b200")
    exe = str(tmp_path / "hdr")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe, "-L", so_dir,
                        "-lvecode_b200", f"-Wl,-rpath,{so_dir}", "-Wl,-rpath-link,/usr/local/cuda/lib64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([exe]).returncode == 0  # vo_version needs no GPU


def test_builtin_cfm_tables_of_the_library_equal_the_oracles():
    """vo_cfm_builtin_table needs no GPU: the literals of src/dat/mod.rs:4, 67-80 as the LIBRARY holds them (csrc/exp.cu) are those of the
    oracle's restatement (oracle/exp_oracle.py), bit for bit — the tables the CFM kernels are driven by."""
    import numpy as np
    import vecode_b200 as vo
    from oracle import exp_oracle as eo
    assert np.array_equal(vo.cfm_table("C_GAUSS_LEGENDRE_4")[0], eo.C_GAUSS_LEGENDRE_4)
    assert np.array_equal(vo.cfm_table("CFM_R2_J1_GL"), eo.CFM_R2_J1_GL)
    assert np.array_equal(vo.cfm_table("CFM_R4_J2_GL"), eo.CFM_R4_J2_GL)
    assert np.array_equal(vo.cfm_table("BLANES17_R4_J4"), eo.BLANES17_R4_J4)


def test_guard_switch_is_read_from_the_environment_without_a_gpu():
    """vo_guard_enabled / vo_guard_check (the library's own bounds check) touch no device while nothing is allocated: off by default, on
    with VECODE_GUARD=1 in the environment of the process that loads the library."""
    import ctypes as C
    import os
    import subprocess
    import sys
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from vecode_b200 import _cabi; l = _cabi.lib(); n = C.c_int64(-1); "
            "print(l.vo_guard_enabled(), l.vo_guard_check(C.byref(n)), n.value)" % ROOT)
    for env_val, want in (("", "0 0 0"), ("1", "1 0 0"), ("0", "0 0 0")):
        env = dict(os.environ)
        env.pop("VECODE_GUARD", None)
        if env_val:
            env["VECODE_GUARD"] = env_val
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0 and r.stdout.split() == want.split(), (env_val, r.stdout, r.stderr)
