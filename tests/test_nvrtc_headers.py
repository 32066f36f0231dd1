"""CPU: the device headers stay NVRTC-compilable (mdb_set_user_potential instantiates the SAME kernel templates with the
user's functor at run time).  nvrtcCompileProgram needs no GPU."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "moleculardynamics.jl_b200", "csrc")


def _nvrtc():
    for name in ("/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    pytest.skip("libnvrtc not found")


def _compile(body, dim=3):
    nv = _nvrtc()
    src = ('#include "kernels.cuh"\nnamespace mdb {\nstruct PotUser {\n    static constexpr bool kSparseHits = true;\n'
           '    __device__ __forceinline__ bool eval(const PotParams &P, double r, double sigma1, double sigma2, double &u, double &f) const\n'
           '    {\n        const double *p = P.p;\n' + body + '\n    }\n'
           '    __device__ __forceinline__ bool may_interact(const PotParams &P, double d2, double, double) const { return d2 < P.p[7]; }\n};\n}\n')
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src.encode(), b"t.cu", 0, None, None) == 0
    names = [b"mdb::k_force_list<%d, mdb::PotUser, true, false>" % dim, b"mdb::k_force_list<%d, mdb::PotUser, false, true>" % dim,
             b"mdb::k_force_overflow<%d, mdb::PotUser, true>" % dim, b"mdb::k_force_cells<%d, mdb::PotUser, true>" % dim,
             b"mdb::k_force_brute<%d, mdb::PotUser, false>" % dim]
    for n in names:
        assert nv.nvrtcAddNameExpression(prog, n) == 0
    opts = [b"-arch=sm_100a", b"-std=c++17", b"-fmad=false", ("-I" + CSRC).encode()]
    rc = nv.nvrtcCompileProgram(prog, len(opts), (C.c_char_p * len(opts))(*opts))
    sz = C.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, C.byref(sz))
    log = C.create_string_buffer(max(sz.value, 1))
    nv.nvrtcGetProgramLog(prog, log)
    lowered = []
    if rc == 0:
        for n in names:
            low = C.c_char_p()
            assert nv.nvrtcGetLoweredName(prog, n, C.byref(low)) == 0
            lowered.append(low.value)
        nv.nvrtcGetCUBINSize(prog, C.byref(sz))
    nv.nvrtcDestroyProgram(C.byref(prog))
    return rc, log.value.decode(), lowered, sz.value


@pytest.mark.parametrize("dim", [2, 3])
def test_user_functor_compiles_to_sm100a_cubin(dim):
    body = ("double s = 0.5 * (sigma1 + sigma2); if (r >= s) { u = 0.0; f = 0.0; return false; }\n"
            "double x = 1.0 - r / s; u = 0.5 * p[0] * x * x; f = p[0] * x / s; return true;")
    rc, log, lowered, size = _compile(body, dim)
    assert rc == 0, log
    assert len(lowered) == 5 and all(b"PotUser" in n for n in lowered) and size > 10000


def test_broken_user_code_reports_a_log():
    rc, log, _, _ = _compile("u = undefined_symbol; return true;")
    assert rc != 0 and "undefined_symbol" in log
