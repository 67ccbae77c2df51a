#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  The one mechanical rewrite needed for g++ to parse OpenCL C: vector literals.

OpenCL C writes a vector literal as a cast of a parenthesised list, `(float3)(a, b, c)`; C++ would parse
that as a cast of a comma expression.  This filter turns each `(T)(...)` with T in {float2, float3,
float4, uchar4} into `T(T_lit{...})` (T_lit: cl_shim.hpp) -- same elements, same order, and C++ evaluates
a braced list left to right, which is the order clang's OpenCL front end uses (it matters at
render.cl:157 and :499, whose elements draw random numbers).  Nothing else is touched.

The rewritten text is written to stdout and piped straight into the compiler (oracle/Makefile, target
_ref); it is never stored: no reference source is copied into this repository.
"""
import re
import sys

CAST = re.compile(r"\(\s*(float2|float3|float4|uchar4)\s*\)\s*\(")


def rewrite(src):
    out = []
    pos = 0
    while True:
        m = CAST.search(src, pos)
        if not m:
            out.append(src[pos:])
            break
        out.append(src[pos:m.start()])
        depth, i = 1, m.end()
        while depth:
            c = src[i]
            depth += (c == "(") - (c == ")")
            i += 1
        # literals nest (`(float4)((float3)(...), 0)` would): rewrite the inside too
        # keep the line breaks a multi-line cast contained (render.cl:115, :498) so line numbers stay put
        t = m.group(1)
        out.append(t + "(" + t + "_lit{" + "\n" * m.group(0).count("\n") + rewrite(src[m.end():i - 1]) + "})")
        pos = i
    return "".join(out)


if __name__ == "__main__":
    text = open(sys.argv[1], encoding="utf-8").read()
    new = rewrite(text)
    # newlines are preserved, so compiler diagnostics keep the reference's line numbers
    assert new.count("\n") == text.count("\n")
    sys.stdout.write('#line 1 "%s"\n' % sys.argv[1])
    sys.stdout.write(new)
