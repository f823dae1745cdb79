"""No Fortran compiler exists in the build image, so the ISO_C_BINDING interface blocks of fortran/*.F90 are
checked against include/zmconv_b200.h with the C compiler instead: from every `bind(C)` interface this test derives
the C prototype the Fortran declarations imply (argument order, scalar-by-value vs array-by-reference, element type,
const-ness from intent(in)) and compiles an assignment of the header's function to a pointer of exactly that type
with -Werror: a wrong argument count, order of types or intent fails the build.  Argument NAMES are compared with the
header's parameter names as well, which catches two same-typed arguments swapped."""
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "zmconv_b200.h")
FILES = ["zm_conv_shim.F90", "zm_conv_intr_batched.F90", "zm_neighbours_shim.F90"]
MIN_INTERFACES = {"zm_neighbours_shim.F90": 4}
CTYPE = {"integer(c_int)": "int", "real(c_double)": "double", "real(c_float)": "float",
         "integer(c_long_long)": "long long", "character(kind=c_char)": "char"}


def _join_continuations(text):
    text = re.sub(r"!.*", "", text)                      # comments
    return re.sub(r"&\s*\n\s*&?", " ", text)


def fortran_interfaces(path):
    """{c_name: (result C type, [(argname, C type)])} of every bind(C) function in the file's interface blocks."""
    src = _join_continuations(open(path).read())
    out = {}
    pat = re.compile(r"(integer\(c_int\)|real\(c_double\))\s+function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)"
                     r"(.*?)end function", re.S | re.I)
    for m in pat.finditer(src):
        res, _, arglist, cname, body = m.groups()
        args = [a.strip() for a in arglist.split(",") if a.strip()]
        decl = {}
        for line in body.splitlines():
            line = line.strip()
            if "::" not in line or line.lower().startswith("import"):
                continue
            left, right = line.split("::", 1)
            attrs = [a.strip().lower() for a in re.split(r",(?![^()]*\))", left)]
            base = attrs[0]
            is_value = "value" in attrs
            intent = next((a for a in attrs if a.startswith("intent")), "")
            for item in re.split(r",(?![^()]*\))", right):
                item = item.strip()
                name = re.match(r"\w+", item).group(0)
                is_array = "(" in item
                if base.startswith("type(c_ptr)"):
                    ct = "void*" if is_value else ("const char**" if name == "names" else "void**")
                elif base.startswith("type("):
                    tname = re.match(r"type\((\w+)\)", base).group(1)
                    ct = ("const " if "intent(in)" in intent else "") + tname + "*"
                else:
                    b = CTYPE[base]
                    if is_value:
                        ct = b
                    else:
                        ct = ("const " if intent == "intent(in)" else "") + b + "*"
                    assert is_value or is_array or intent in ("intent(inout)", "intent(out)", "intent(in)"), (cname, name)
                decl[name.lower()] = ct
        out[cname] = (CTYPE[res.lower()], [(a, decl[a.lower()]) for a in args])
    return out


def header_param_names():
    hdr = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    names = {}
    for m in re.finditer(r"\b(?:int|void|double|long long)\s+(zm_\w+)\s*\(([^;]*?)\)\s*;", hdr, re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        names[m.group(1)] = [re.search(r"(\w+)\s*$", p).group(1) for p in params if p and p != "void"]
    return names


@pytest.mark.parametrize("fname", FILES)
def test_bind_c_interfaces_match_the_header(fname):
    ifs = fortran_interfaces(os.path.join(ROOT, "fortran", fname))
    assert len(ifs) >= MIN_INTERFACES.get(fname, 6), sorted(ifs)
    hnames = header_param_names()
    lines = ['#include "zmconv_b200.h"']
    for cname, (res, args) in sorted(ifs.items()):
        assert cname in hnames, f"{fname}: {cname} is not declared in include/zmconv_b200.h"
        got = [a.lower() for a, _ in args]
        want = [n.lower() for n in hnames[cname]]
        # the header names a few parameters differently from the Fortran dummies; compare position by position
        alias = {"p": "p", "buf": "buf", "on": "on"}
        assert len(got) == len(want), (cname, len(got), len(want))
        for g, w in zip(got, want):
            assert g == w or alias.get(g) == w or g.rstrip("_") == w or w.startswith(g) or g.startswith(w), (cname, g, w)
        proto = ", ".join(t for _, t in args) or "void"
        lines.append(f"static {res} (*const chk_{cname})({proto}) = &{cname};")
    lines.append("int main(void) { return 0; }")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "chk.c")
        open(src, "w").write("\n".join(lines) + "\n")
        r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-Wno-unused-const-variable", "-Wno-unused-variable",
                            "-I", os.path.join(ROOT, "include"), "-c", src, "-o", os.path.join(td, "chk.o")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + "\n" + "\n".join(lines)


def test_the_check_detects_a_swapped_or_missing_argument():
    """The generated assignment must fail to compile when the prototype is wrong (guards the test itself)."""
    bad = ['#include "zmconv_b200.h"',
           "static int (*const chk)(int, const double*, const int*, int, const double*, const double*, double*, double, const int*) = &zm_conv_tend_2_batch;",
           "int main(void) { return 0; }"]
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "bad.c")
        open(src, "w").write("\n".join(bad) + "\n")
        r = subprocess.run(["gcc", "-std=c11", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", src, "-o",
                            os.path.join(td, "bad.o")], capture_output=True, text=True)
        assert r.returncode != 0
