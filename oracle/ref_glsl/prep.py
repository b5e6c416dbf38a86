#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Mechanical GLSL -> C++ literal transform for oracle/_ref (never touches the semantics of a statement).

GLSL floating literals without suffix are 32-bit (`1.0 - x` is an fp32 subtraction); the same token in C++ is a double and would
promote the whole expression. The only edits made to the reference's shader text are therefore:
  1. every floating literal without a suffix gets an `f` suffix,
  2. `#extension ...` lines (not a C++ directive) are dropped.
Usage: prep.py <in.glsl> <out.glsl> [first_line last_line]   (the optional range extracts a function from a shader that also holds
layout() declarations C++ cannot parse)."""
import re
import sys

LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")


def transform(text):
    out = []
    for line in text.splitlines():
        if line.lstrip().startswith("#extension"):
            continue
        if line.lstrip().startswith("#include"):
            out.append(line)
            continue
        code, sep, comment = line.partition("//")
        out.append(LIT.sub(lambda m: m.group(1) + "f", code) + sep + comment)
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    src = open(sys.argv[1]).read()
    if len(sys.argv) > 3:
        a, b = int(sys.argv[3]), int(sys.argv[4])
        src = "\n".join(src.splitlines()[a - 1:b]) + "\n"
    open(sys.argv[2], "w").write(transform(src))
