#!/usr/bin/env python3
"""oracle/glsl_cpu/gen_shader_cpp.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Rewrites GLSL 4.30 compute-shader sources of the REFERENCE (read from where they lie under /root/reference) into text
that a C++ compiler accepts as the body of a struct, for oracle/glsl_cpu/glsl_emu.h.  Only declarations and qualifiers
are touched; statements and expressions go through unchanged:

  #version, BOM, comments                 removed
  layout(local_size_*) in;                removed
  [layout(...)] uniform T name[N];        T name = U<T>("name");  /  arr<T,N> name = UA<T,N>("name");   (once per name)
  out / inout / in parameter qualifiers   T& / T& / T
  float literals without a suffix         get an 'f' (GLSL literals are fp32; C++ ones would be double)
  function prototypes                     removed (the linked shader files become one struct body)
  a struct defined by several files       kept once

The output goes to a build directory outside the repository (the Makefile passes a temporary one): no reference
source is copied into the tree.  usage: gen_shader_cpp.py OUT.h Program file1.comp [file2.comp ...]
"""
import re
import sys


def strip_comments(src):
    out, i, n = [], 0, len(src)
    while i < n:
        if src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif src.startswith("/*", i):
            j = src.find("*/", i + 2)
            j = n if j < 0 else j + 2
            out.append("\n" * src.count("\n", i, j))
            i = j
        else:
            out.append(src[i])
            i += 1
    return "".join(out)


FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def top_level_items(src):
    """Split into items at brace depth 0: preprocessor lines, declarations ending in ';', blocks ending in '}'."""
    items, cur, depth, i, n = [], [], 0, 0, len(src)
    at_line_start = True
    while i < n:
        ch = src[i]
        if depth == 0 and at_line_start and ch == "#":
            j = src.find("\n", i)
            j = n if j < 0 else j + 1
            if "".join(cur).strip():
                items.append("".join(cur))
            else:
                items.append("".join(cur))      # keep blank lines so that line numbers stay close
            cur = []
            items.append(src[i:j])
            i = j
            at_line_start = True
            continue
        cur.append(ch)
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                # a struct definition continues up to its ';', a function body ends here
                k = i + 1
                while k < n and src[k] in " \t\r\n":
                    k += 1
                if k < n and src[k] == ";" and re.match(r"\s*struct\b", "".join(cur)):
                    cur.append(src[i + 1:k + 1])
                    i = k
                items.append("".join(cur))
                cur = []
        elif ch == ";" and depth == 0:
            items.append("".join(cur))
            cur = []
        if ch == "\n":
            at_line_start = True
        elif ch not in " \t\r":
            at_line_start = False
        i += 1
    if "".join(cur).strip():
        items.append("".join(cur))
    return items


PROTOTYPE = re.compile(r"^\s*[A-Za-z_]\w*\s+[A-Za-z_]\w*\s*\([^;{}]*\)\s*;\s*$", re.S)
UNIFORM = re.compile(r"^(\s*)(?:layout\s*\([^)]*\)\s*)?uniform\s+(\w+)\s+(\w+)\s*(?:\[\s*(\d+)\s*\])?\s*;\s*$", re.S)
LAYOUT_IN = re.compile(r"^\s*layout\s*\([^)]*\)\s*in\s*;\s*$", re.S)
STRUCT = re.compile(r"^\s*struct\s+(\w+)")


def convert(paths):
    seen_uniforms, seen_structs, out = set(), set(), []
    for path in paths:
        src = open(path, encoding="utf-8-sig").read().replace("\r\n", "\n")
        src = strip_comments(src)
        src = re.sub(r"^[ \t]*#version[^\n]*", "", src, flags=re.M)
        src = FLOAT_LIT.sub(r"\1f", src)
        src = re.sub(r"\b(?:out|inout)\s+(\w+)\s+(\w+)", r"\1& \2", src)
        src = re.sub(r"(?<=[(,])(\s*)in\s+(\w+\s+\w+)", r"\1\2", src)
        out.append('#line 1 "%s"\n' % path)
        for item in top_level_items(src):
            keep_lines = "\n" * item.count("\n")
            if LAYOUT_IN.match(item) or PROTOTYPE.match(item):
                out.append(keep_lines)
                continue
            m = UNIFORM.match(item)
            if m:
                lead, typ, name, count = m.groups()
                if name in seen_uniforms:
                    out.append(keep_lines)
                    continue
                seen_uniforms.add(name)
                nl = "\n" * (item.count("\n") - lead.count("\n"))
                if count:
                    out.append('%sarr<%s,%s> %s = UA<%s,%s>("%s");%s' % (lead, typ, count, name, typ, count, name, nl))
                else:
                    out.append('%s%s %s = U<%s>("%s");%s' % (lead, typ, name, typ, name, nl))
                continue
            m = STRUCT.match(item)
            if m and "{" in item:
                if m.group(1) in seen_structs:
                    out.append(keep_lines)
                    continue
                seen_structs.add(m.group(1))
            out.append(item)
        out.append("\n")
    return "".join(out)


def main():
    out_path, struct_name, paths = sys.argv[1], sys.argv[2], sys.argv[3:]
    body = convert(paths)
    with open(out_path, "w") as f:
        f.write("// generated by oracle/glsl_cpu/gen_shader_cpp.py from the reference's shader files; not kept\n")
        f.write("struct %s : glsl::Invocation {\n%s\n};\n" % (struct_name, body))


if __name__ == "__main__":
    main()
