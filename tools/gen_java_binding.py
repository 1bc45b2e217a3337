#!/usr/bin/env python
"""Generates integration/java/eu/heros/gpu/HerosGpuNative.java from include/heros_gpu.h: one java.lang.foreign downcall
handle per entry point of the C ABI, with the FunctionDescriptor that matches its C prototype (the file a maintainer of
the reference adds to bind libheros_gpu.so through the Panama FFM API; INTEGRATION.md shows the classes built on it).

There is no JDK in this image, so the file is not compiled here; tests/test_abi_and_host.py checks that it is up to date
with the header (every function, the right arity and value layouts).

usage: python tools/gen_java_binding.py [--check]
"""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "heros_gpu.h")
OUT = os.path.join(ROOT, "integration", "java", "eu", "heros", "gpu", "HerosGpuNative.java")

SCALARS = {"int8_t": "JAVA_BYTE", "uint8_t": "JAVA_BYTE", "int16_t": "JAVA_SHORT", "uint16_t": "JAVA_SHORT",
           "int32_t": "JAVA_INT", "uint32_t": "JAVA_INT", "int64_t": "JAVA_LONG", "uint64_t": "JAVA_LONG",
           "double": "JAVA_DOUBLE", "float": "JAVA_FLOAT"}


def prototypes(text: str):
    """(return type, name, [(type, name)]) of every function the header declares"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for ret, name, args in re.findall(r"^\s*((?:const\s+)?[a-z_0-9]+\s*\*?)\s*(heros_[a-z_]+)\s*\(([^;{]*?)\)\s*;", text,
                                      flags=re.M | re.S):
        params = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"(.*?)([A-Za-z_][A-Za-z_0-9]*)(\[\d*\])?$", a)
                ctype, pname, array = m.group(1).strip(), m.group(2), m.group(3)
                if not ctype:                      # unnamed parameter: "const float*"
                    ctype, pname = a, "arg%d" % len(params)
                if array:
                    ctype += "*"
                params.append((ctype, pname))
        out.append((" ".join(ret.split()), name, params))
    return out


def constants(text: str):
    """name -> value of the #define / enum integer constants of the header (array bounds of the structs)"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    values = {}
    for name, val in re.findall(r"#define\s+(HEROS_[A-Z_0-9]+)\s+(\d+)u?\b", text):
        values[name] = int(val)
    for body in re.findall(r"enum\s*[a-z_]*\s*\{(.*?)\}", text, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = (x.strip() for x in item.split("="))
                nxt = int(val, 0) if re.fullmatch(r"-?\w+", val) and val.lstrip("-").isdigit() else values[val]
            else:
                name = item
            values[name] = nxt
            nxt += 1
    return values


def structs(text: str):
    """name -> [(field, base C type, [array dimensions])] of the plain-old-data structs of the header"""
    consts = constants(text)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s+\w+\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            ctype, rest = decl.split(" ", 1)
            for item in rest.split(","):
                m = re.match(r"\s*(\w+)((?:\[\w+\])*)\s*$", item)
                dims = [consts[d] if d in consts else int(d) for d in re.findall(r"\[(\w+)\]", m.group(2))]
                fields.append((m.group(1), ctype, dims))
        out[name] = fields
    return out


SIZES = {"JAVA_BYTE": 1, "JAVA_SHORT": 2, "JAVA_INT": 4, "JAVA_LONG": 8, "JAVA_DOUBLE": 8, "JAVA_FLOAT": 4}


def struct_size(name: str, all_structs) -> int:
    total = 0
    for _, ctype, dims in all_structs[name]:
        one = struct_size(ctype, all_structs) if ctype in all_structs else SIZES[SCALARS[ctype]]
        for d in dims:
            one *= d
        total += one
    return total


def render_struct(name: str, all_structs) -> list:
    parts = []
    for field, ctype, dims in all_structs[name]:
        elem = f"{ctype.upper()}_LAYOUT" if ctype in all_structs else SCALARS[ctype]
        for d in reversed(dims):
            elem = f"MemoryLayout.sequenceLayout({d}, {elem})"
        parts.append(f'{elem}.withName("{field}")')
    lines = [f"    /** {name}: {struct_size(name, all_structs)} bytes, natural alignment, no implicit padding */",
             f"    public static final StructLayout {name.upper()}_LAYOUT = MemoryLayout.structLayout("]
    for i, part in enumerate(parts):
        lines.append(f"            {part}{',' if i + 1 < len(parts) else ''}")
    lines[-1] += ");"
    lines.append("")
    return lines


def layout(ctype: str) -> str:
    if "*" in ctype or ctype.replace("const", "").strip() == "heros_handle":
        return "ADDRESS"
    base = ctype.replace("const", "").strip()
    if base not in SCALARS:
        raise ValueError(f"no value layout for C type {ctype!r}")
    return SCALARS[base]


def constant(name: str) -> str:
    return name[len("heros_"):].upper()


def render(protos, all_structs=None) -> str:
    lines = ["// GENERATED by tools/gen_java_binding.py from include/heros_gpu.h -- do not edit.",
             "package eu.heros.gpu;", "",
             "import java.lang.foreign.Arena;", "import java.lang.foreign.FunctionDescriptor;",
             "import java.lang.foreign.Linker;", "import java.lang.foreign.MemoryLayout;", "import java.lang.foreign.StructLayout;",
             "import java.lang.foreign.SymbolLookup;",
             "import java.lang.invoke.MethodHandle;", "",
             "import static java.lang.foreign.ValueLayout.ADDRESS;", "import static java.lang.foreign.ValueLayout.JAVA_BYTE;",
             "import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;", "import static java.lang.foreign.ValueLayout.JAVA_FLOAT;",
             "import static java.lang.foreign.ValueLayout.JAVA_INT;", "import static java.lang.foreign.ValueLayout.JAVA_LONG;",
             "import static java.lang.foreign.ValueLayout.JAVA_SHORT;", "",
             "/**", " * Downcall handles for every entry point of libheros_gpu.so (include/heros_gpu.h), JDK 22+ (java.lang.foreign).",
             " * Status codes, struct layouts and call order are documented in the header; INTEGRATION.md shows the plug-in",
             " * classes (GpuCovid19Transmission, GpuCovid19Progression) built on these handles.",
             " */", "public final class HerosGpuNative", "{",
             "    private static final Linker LINKER = Linker.nativeLinker();",
             "    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(\"libheros_gpu.so\", Arena.global());", "",
             "    private static MethodHandle fn(final String name, final FunctionDescriptor descriptor)", "    {",
             "        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), descriptor);", "    }", "",
             "    private HerosGpuNative()", "    {", "    }", ""]
    for name in (all_structs or {}):
        lines += render_struct(name, all_structs)
    for ret, name, params in protos:
        sig = ", ".join(f"{t} {n}" for t, n in params) or "void"
        layouts = [layout(ret)] + [layout(t) for t, _ in params]
        lines.append(f"    /** {ret} {name}({sig}) */")
        lines.append(f"    public static final MethodHandle {constant(name)} = fn(\"{name}\",")
        lines.append(f"            FunctionDescriptor.of({', '.join(layouts)}));")
        lines.append("")
    lines[-1] = "}"
    return "\n".join(lines) + "\n"


def main():
    with open(HEADER) as f:
        header = f.read()
    text = render(prototypes(header), structs(header))
    if "--check" in sys.argv:
        with open(OUT) as f:
            if f.read() != text:
                print("integration/java/.../HerosGpuNative.java is out of date: run tools/gen_java_binding.py")
                return 1
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write(text)
    print(OUT, len(prototypes(open(HEADER).read())), "entry points")
    return 0


if __name__ == "__main__":
    sys.exit(main())
