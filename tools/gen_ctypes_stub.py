#!/usr/bin/env python
"""Generate the ctypes struct definitions of INTEGRATION.md from include/tigar_b200.h, so the
stub a maintainer copies can never lag behind the header.

    python tools/gen_ctypes_stub.py            # prints the stub
    python tools/gen_ctypes_stub.py --write    # rewrites the block between the markers in INTEGRATION.md
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BEGIN, END = "<!-- BEGIN GENERATED STRUCTS -->", "<!-- END GENERATED STRUCTS -->"
CT = {"int32_t": "C.c_int32", "int64_t": "C.c_int64", "double": "C.c_double"}


def parse_structs(header_text):
    """-> {name: [(field, ctype string)]} for the typedef structs of the header."""
    text = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S)
    dims = dict(re.findall(r"#define\s+(TG_\w+)\s+(\d+)", text))
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(\w+)(\[(\w+)\])?$", decl)
            if m is None:
                raise ValueError("cannot parse field %r of %s" % (decl, name))
            base, ptr, fname, n = m.group(2), m.group(3), m.group(4), m.group(6)
            ct = "C.c_void_p" if ptr else CT[base]
            if n:
                ct += " * %s" % dims.get(n, n)
            fields.append((fname, ct))
        out[name] = fields
    return out


def stub(structs):
    L = []
    for name in ("tg_basis", "tg_win"):
        L.append("class %s(C.Structure):        # include/tigar_b200.h: %s" % (name, name))
        L.append("    _fields_ = [")
        for f, ct in structs[name]:
            L.append('        ("%s", %s),' % (f, ct))
        L.append("    ]")
        L.append("")
    L.append("assert lib.tg_sizeof_win() == C.sizeof(tg_win)")
    L.append("assert lib.tg_sizeof_basis() == C.sizeof(tg_basis)")
    return "\n".join(L)


def main():
    hdr = open(os.path.join(ROOT, "include", "tigar_b200.h")).read()
    text = stub(parse_structs(hdr))
    if "--write" not in sys.argv:
        print(text)
        return
    path = os.path.join(ROOT, "INTEGRATION.md")
    md = open(path).read()
    a, b = md.index(BEGIN) + len(BEGIN), md.index(END)
    md = md[:a] + "\n```python\n" + text + "\n```\n" + md[b:]
    open(path, "w").write(md)


if __name__ == "__main__":
    main()
