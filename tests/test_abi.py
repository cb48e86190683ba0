"""CPU checks of the drop-in boundary: the C-ABI library exists, loads, and exports every symbol
include/ufair.h declares (no compute calls -- those need a GPU)."""
import ctypes
import os
import re

import pytest

from fiveeqscm_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ufair.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ufair_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_abi.lib_path()):
        import __graft_entry__ as g
        g.build()
    return _abi.lib_path()


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("ufair_run_f64", "ufair_run_f32", "ufair_hfc_pulse_f64", "ufair_stats_finalize",
                 "ufair_g1g0_f64", "ufair_kq_f64", "ufair_run_host_f64", "ufair_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built)
    for name in _declared_functions():
        assert hasattr(L, name), f"{name} declared in include/ufair.h but not exported by libufair.so"


def test_binding_covers_every_declared_symbol():
    assert sorted(_abi.SIGNATURES) == _declared_functions()


def test_abi_version_and_struct_size(built):
    L = _abi.lib()
    assert L.ufair_abi_version() == _abi.ABI_VERSION
    assert L.ufair_block_members() % 32 == 0 and L.ufair_block_members() > 0
    # a descriptor with the wrong struct_size is rejected before anything touches the device
    d = _abi.UfairDesc(n_gas=1, n_t=1, n_member=1, ld_member=2)
    d.struct_size = 8
    assert L.ufair_run_f64(ctypes.byref(d), None) == _abi.ERR_ARG
    assert b"struct_size" in L.ufair_last_error()


def test_argument_errors_are_reported_not_fatal(built):
    L = _abi.lib()
    d = _abi.UfairDesc(n_gas=9, n_t=1, n_member=1, ld_member=2)
    assert L.ufair_run_f64(ctypes.byref(d), None) == _abi.ERR_ARG
    d = _abi.UfairDesc(n_gas=1, n_t=4, n_member=3, ld_member=3)   # 24-byte rows: not 16-byte aligned
    assert L.ufair_run_f64(ctypes.byref(d), None) == _abi.ERR_ALIGN
    d = _abi.UfairDesc(n_gas=1, n_t=4, n_member=4, ld_member=4)   # NULL inputs
    assert L.ufair_run_f64(ctypes.byref(d), None) == _abi.ERR_ARG
    d = _abi.UfairDesc(n_gas=1, n_t=4, n_member=4, ld_member=4, alpha_mode=7)
    assert L.ufair_run_f64(ctypes.byref(d), None) == _abi.ERR_ARG
    d = _abi.UfairDesc(n_gas=1, n_t=4, n_member=4, ld_member=2 ** 31)   # TMA coordinates are 32-bit
    assert L.ufair_run_f64(ctypes.byref(d), None) == _abi.ERR_ARG and b"split the member axis" in L.ufair_last_error()
    with pytest.raises(_abi.UfairError):
        _abi.check(L.ufair_run_f32(ctypes.byref(d), None))


def _c_layout(header, struct, fields, tmp_path):
    """sizeof(struct) and offsetof every field, as the C compiler lays the header's struct out."""
    import subprocess
    src = tmp_path / f"layout_{struct}.c"
    exe = tmp_path / f"layout_{struct}"
    body = "".join(f'  printf("{f} %zu\\n", offsetof({struct}, {f}));\n' for f in fields)
    src.write_text(f'#include <stddef.h>\n#include <stdio.h>\n#include "{header}"\nint main(void) {{\n'
                   f'  printf("sizeof %zu\\n", sizeof({struct}));\n{body}  return 0;\n}}\n')
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    return dict(zip(out[0::2], (int(v) for v in out[1::2])))


@pytest.mark.parametrize("which", ["ufair_desc", "ufair_sampler", "ufo_desc"])
def test_ctypes_mirrors_match_the_c_headers_field_by_field(which, tmp_path):
    """Each binding is checked against ITS OWN header by the C compiler (offsetof every field), so the
    product's descriptor (include/ufair.h <-> _abi.UfairDesc) and the oracle's (oracle/ufo.h <->
    c_oracle.UfoDesc) are each right on their own -- they share no code that could be wrong twice."""
    from oracle import c_oracle
    header, cls = {"ufair_desc": (HEADER, _abi.UfairDesc), "ufair_sampler": (HEADER, _abi.UfairSampler),
                   "ufo_desc": (os.path.join(ROOT, "oracle", "ufo.h"), c_oracle.UfoDesc)}[which]
    names = [f[0] for f in cls._fields_]
    lay = _c_layout(header, which, names, tmp_path)
    assert lay["sizeof"] == ctypes.sizeof(cls)
    for n in names:
        assert lay[n] == getattr(cls, n).offset, f"{which}.{n}: C offset {lay[n]} != ctypes {getattr(cls, n).offset}"
    # and the header has no field the binding forgot
    src = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (which, which), src, flags=re.S).group(1)
    declared = [re.search(r"(\w+)\s*(?:\[[^\]]*\])*\s*$", part.strip()).group(1)
                for stmt in body.split(";") if stmt.strip() for part in stmt.split(",")]
    assert sorted(declared) == sorted(names)


def test_oracle_descriptor_is_its_own(built):
    """oracle/c_oracle.py must not borrow the product's binding, and the C oracle must not include the
    product's header (a layout mistake would be common-mode)."""
    from oracle import c_oracle
    assert "fiveeqscm_b200" not in open(os.path.join(ROOT, "oracle", "c_oracle.py")).read().replace(
        "fiveeqscm_b200/_abi.py", "")
    for f in ("ufair_oracle.c", "ufair_oracle_fast.c", "ufo.h"):
        assert "ufair.h" not in re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "oracle", f)).read(), flags=re.S)
    d = c_oracle.UfoDesc(n_gas=1, n_t=0, n_member=0, ld_member=0)
    assert c_oracle.lib().ufo_run_f64(ctypes.byref(d), 1) >= 1
    d.struct_size = 8
    assert c_oracle.lib().ufo_run_f64(ctypes.byref(d), 1) < 0


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product packages may import, load or
    link it (no `import oracle`, no libufair_oracle, no ufo_* symbol)."""
    for pkg in ("fiveeqscm_b200", "U_FaIR", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            if os.path.basename(dirpath).startswith("build"):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f)).read()
                    for needle in ("import oracle", "from oracle", "libufair_oracle", "ufo_run", "c_oracle"):
                        assert needle not in text, f"{os.path.join(dirpath, f)} mentions {needle!r}"
