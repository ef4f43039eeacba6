"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/trpo_b200.h declares,
keeps the reference's TRPOparam layout, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "trpo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n.startswith(("trpo_", "FVP_", "CG_", "TRPO_"))))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for must in ("FVP_GPU", "CG_GPU", "TRPO_Update_GPU", "FVP_FPGA", "CG_FPGA", "trpo_ctx_create", "trpo_ctx_fvp",
                 "trpo_ctx_cg", "trpo_ctx_update", "trpo_ctx_init_comm", "TRPO_Lightweight_GPU", "TRPO_Lightweight_FPGA",
                 "trpo_ctx_rollout_arm", "trpo_ctx_set_rollout", "trpo_vf_advantage", "trpo_vf_evaluate",
                 "trpo_batch_file_write", "trpo_ctx_set_batch_file"):
        assert must in names


def test_library_exports_every_declared_symbol(pkg):
    lib = C.CDLL(pkg.api.library_path())
    dropin = C.CDLL(pkg.api.library_path(dropin=True))
    for name in declared_functions():
        target = dropin if name in ("FVP_FPGA", "CG_FPGA", "TRPO_Lightweight_FPGA") else lib
        assert hasattr(target, name), f"{name} declared in include/trpo_b200.h but not exported"
    # the FPGA names must NOT leak into the plain library (they would clash with a MaxCompiler build)
    with pytest.raises(AttributeError):
        lib.FVP_FPGA


def test_trpoparam_layout_matches_reference(pkg):
    # /root/reference/src/include/TRPO.h:6-49 -- 11 fields x 8 bytes on LP64, in this order
    P = pkg.TRPOparam
    assert C.sizeof(P) == 88
    names = [f[0] for f in P._fields_]
    assert names == ["ModelFile", "BaselineFile", "ResultFile", "DataFile", "NumLayers", "AcFunc", "LayerSize",
                     "NumSamples", "CG_Damping", "PaddedLayerSize", "NumBlocks"]
    assert P.NumSamples.offset == 56 and P.CG_Damping.offset == 64


def test_num_params_matches_reference_formula(pkg):
    assert pkg.api.num_params([15, 16, 16, 3]) == 582          # ArmTest
    assert pkg.api.num_params([17, 64, 64, 6]) == 5708
    assert pkg.api.num_params([376, 256, 256, 17]) == 166690
    assert pkg.api.num_params([4, 5, 2]) == pkg.synth.num_params([4, 5, 2])


def test_missing_files_return_minus_one_like_the_reference(pkg, tmp_path, capfd):
    # TRPO_FVP_FPGA.c:102-105: "[ERROR] Cannot open Model File [...]" then return -1
    v = np.zeros(582)
    out, t = pkg.FVP_GPU(str(tmp_path / "nope_model.txt"), str(tmp_path / "nope_data.txt"), [15, 16, 16, 3], "lttl", 10, 0.1, v)
    assert t == -1.0
    assert "[ERROR] Cannot open Model File" in capfd.readouterr().err
    mf = tmp_path / "m.txt"
    pkg.textio.write_model(str(mf), np.zeros(582))
    out, t = pkg.CG_GPU(str(mf), str(tmp_path / "nope_data.txt"), [15, 16, 16, 3], "lttl", 10, 0.1, v)
    assert t == -1.0
    assert "[ERROR] Cannot open Data File" in capfd.readouterr().err


def test_unsupported_activation_is_refused(pkg, capfd):
    with pytest.raises(RuntimeError):
        pkg.Context([4, 5, 2], "lxl")
    assert "Unsupported" in capfd.readouterr().err


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.Context([15, 16, 16, 3], "lttl")


def test_product_does_not_reference_the_oracle():
    """The product sources must never include, link or dlopen anything under oracle/."""
    pkg_dir = os.path.join(ROOT, "trpo-robot-control_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".cu", ".cuh", ".c", ".h", ".py")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "trpo_oracle" not in text and "oracle_lib" not in text, f


def test_header_compiles_as_c_and_cxx_and_after_the_reference_header(tmp_path):
    """include/trpo_b200.h is plain C (gnu11 and c99), has C linkage under C++, and can follow the reference's own TRPO.h
    (it then reuses that TRPOparam instead of redefining it)."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "t.c"
    src.write_text('#include "trpo_b200.h"\nint main(void) { TRPOparam p; (void)p; return (int)sizeof(trpo_batch_file_header) - 64; }\n')
    for cmd in (["gcc", "-std=gnu11", "-Wall", "-Werror"], ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror"],
                ["g++", "-std=c++17", "-x", "c++", "-Wall", "-Werror"]):
        r = subprocess.run(cmd + ["-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, (cmd, r.stderr)
    ref_inc = "/root/reference/src/include"
    if os.path.exists(os.path.join(ref_inc, "TRPO.h")):
        src2 = tmp_path / "t2.c"
        src2.write_text('#include "TRPO.h"\n#include "trpo_b200.h"\n'
                        'double f(TRPOparam p, double *r, double *b) { return CG_GPU(p, r, b, 10, 1e-10, 1) + CG_FPGA(p, r, b, 10, 1e-10, 1); }\n')
        r = subprocess.run(["gcc", "-std=gnu11", "-w", "-fsyntax-only", "-I", ref_inc, "-I", inc, str(src2)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_reference_style_caller_links_against_the_dropin_library(pkg, tmp_path):
    """The link line of INTEGRATION.md section 1: a C caller written against the reference's prototypes (TRPO.h:98,101,113)
    links to libtrpo_b200_dropin.so with no other glue."""
    import subprocess
    libdir = os.path.dirname(pkg.api.library_path())
    src = tmp_path / "caller.c"
    src.write_text(
        '#include <stddef.h>\n'
        'typedef struct { char *ModelFile, *BaselineFile, *ResultFile, *DataFile; size_t NumLayers; char *AcFunc; size_t *LayerSize;\n'
        '                 size_t NumSamples; double CG_Damping; size_t *PaddedLayerSize, *NumBlocks; } TRPOparam;\n'
        'double FVP_FPGA(TRPOparam param, double *Result, double *Input);\n'
        'double CG_FPGA(TRPOparam param, double *Result, double *b, size_t MaxIter, double ResidualTh, size_t NumThreads);\n'
        'double TRPO_Lightweight_FPGA(TRPOparam param, const int NumIter, const size_t NumThreads);\n'
        'int main(int argc, char **argv) { TRPOparam p = {0}; double r[4], b[4];\n'
        '  if (argc > 99) return (int)(FVP_FPGA(p, r, b) + CG_FPGA(p, r, b, 10, 1e-10, 1) + TRPO_Lightweight_FPGA(p, 1, 1));\n'
        '  return 0; }\n')
    exe = tmp_path / "caller"
    r = subprocess.run(["gcc", "-std=gnu11", str(src), "-L", libdir, "-ltrpo_b200_dropin", f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)], capture_output=True).returncode == 0      # loads (CUDA runtime and all) and exits
