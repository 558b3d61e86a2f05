"""CPU: the C-ABI library loads and exports every symbol include/ldmseg_b200.h declares (no compute calls)."""
import os
import re

from video_latent_diffusion_panoptic_segmentation_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ldmseg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ldmseg_b200.h but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(L.SIGNATURES) == names


def test_abi_version_and_error_string():
    lib = L.load()
    assert lib.ldm_abi_version() == 4  # 4: LayerNorm-fold fields; 3: splitk_ws; 2: qkv_part0 / kv_seq descriptor fields
    assert isinstance(lib.ldm_last_error(), bytes)
    assert lib.ldm_launch_count() >= 0


def test_struct_sizes_match_header_layout():
    import ctypes as C
    # 12 pointers + float + 14 int32, 8-byte aligned
    assert C.sizeof(L.GemmDesc) == 12 * 8 + 4 + 14 * 4 + 0 or C.sizeof(L.GemmDesc) % 8 == 0
    assert C.sizeof(L.AttnDesc) == 4 * 8 + 6 * 4 + 4 + 4 + 4 + 4  # kv_seq + tail padding to 8 bytes
    assert C.sizeof(L.GroupNormDesc) == 6 * 8 + 5 * 4 + 4 + 4 + 4


def test_ctypes_signatures_have_the_declared_arity():
    """Every prototype of include/ldmseg_b200.h against the ctypes argtypes of _lib.py: same number of parameters, and
    pointer / integer / floating parameters in the same positions (catches a binding that drifted from the header)."""
    import ctypes as C
    src = open(os.path.join(ROOT, "include", "ldmseg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|long long|const char\*)\s+(ldm_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert len(protos) == len(L.SIGNATURES)
    for name, params in protos:
        params = [p.strip() for p in params.split(",")] if params.strip() not in ("", "void") else []
        argtypes = L.SIGNATURES[name][1]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, ctypes {len(argtypes)}"
        for prm, ct in zip(params, argtypes):
            is_ptr = "*" in prm or prm.startswith("ldm_stream_t")
            if is_ptr:
                assert ct is C.c_void_p or isinstance(ct, type) and issubclass(ct, C._Pointer), (name, prm, ct)
            elif prm.startswith(("float", "double")):
                assert ct in (C.c_float, C.c_double), (name, prm, ct)
                assert (ct is C.c_double) == prm.startswith("double"), (name, prm, ct)
            else:
                assert ct in (C.c_int, C.c_int32, C.c_int64, C.c_longlong), (name, prm, ct)
                if prm.startswith("int64_t"):
                    assert C.sizeof(ct) == 8, (name, prm, ct)


def test_descriptor_structs_list_the_header_fields_in_order():
    """ldm_gemm_desc / ldm_attn_desc / ldm_groupnorm_desc: field names and order of the ctypes Structures equal the
    header's (a reordered or missing field would silently shift every later one)."""
    src = open(os.path.join(ROOT, "include", "ldmseg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for cname, ctype in (("ldm_gemm_desc", L.GemmDesc), ("ldm_attn_desc", L.AttnDesc),
                         ("ldm_groupnorm_desc", L.GroupNormDesc)):
        body = re.search(r"typedef struct " + cname + r"\s*\{(.*?)\}\s*" + cname + r"\s*;", src, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const void* a1" / "int32_t B, H, W" / "float ln_eps": the identifiers after the type
            first, *rest = decl.split(",")
            names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", first)[-1])
            names += [r.strip().lstrip("*") for r in rest]
        assert names == [f[0] for f in ctype._fields_], (cname, names)


def test_header_compiles_as_c_and_struct_offsets_match_ctypes(tmp_path):
    """include/ldmseg_b200.h is plain C (gcc -std=c99 -pedantic-errors) and the offsets / sizes the C compiler gives the
    three descriptor structs are the ones ctypes uses."""
    import ctypes as C
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    structs = (("ldm_gemm_desc", L.GemmDesc), ("ldm_attn_desc", L.AttnDesc), ("ldm_groupnorm_desc", L.GroupNormDesc))
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "ldmseg_b200.h"', "int main(void) {"]
    for cname, ctype in structs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in ctype._fields_:
            lines.append(f'  printf("{cname}.{f[0]} %zu\\n", offsetof({cname}, {f[0]}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-pedantic-errors", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True, capture_output=True, text=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ctype in structs:
        assert int(out[cname]) == C.sizeof(ctype), cname
        for f in ctype._fields_:
            assert int(out[f"{cname}.{f[0]}"]) == getattr(ctype, f[0]).offset, (cname, f[0])


def test_product_library_reads_no_environment():
    """The header promises no global mutable state: the A/B switches (LDM_GEMM_*, LDM_ATTN_*, LDM_GN_*, LDM_PDL) exist only
    in diagnostic builds (-DLDM_DIAG, LDM_BUILD_DIAG=1); the product .so does not even import getenv."""
    import subprocess
    from video_latent_diffusion_panoptic_segmentation_b200 import _lib as L
    from video_latent_diffusion_panoptic_segmentation_b200 import build as B
    if B.DIAG:
        import pytest
        pytest.skip("diagnostic build requested through LDM_BUILD_DIAG")
    out = subprocess.run(["nm", "-D", "--undefined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "getenv" not in out, "libldmseg_b200.so imports getenv: a diagnostic switch leaked into the product build"
    strings = subprocess.run(["strings", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "LDM_GEMM_SPLITK" not in strings and "LDM_ATTN_POLY" not in strings
