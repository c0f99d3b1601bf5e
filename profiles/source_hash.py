"""Hash of the CUDA sources the library is built from, comments and white space removed — what an ncu profile has to match to
describe a build (bench.py's roofline.profile_matches_build; profiles/derive_fused_json.py stamps it). A comment edit does not
change it, any token does. usage: python profiles/source_hash.py [csrc directory]"""
import hashlib
import os
import sys


def strip_code(text: str) -> str:
    """C/C++ source without comments and without white space (string and character literals kept verbatim)."""
    out, i, n = [], 0, len(text)
    while i < n:
        c = text[i]
        if c == "/" and i + 1 < n and text[i + 1] == "/":
            while i < n and text[i] != "\n":
                i += 1
        elif c == "/" and i + 1 < n and text[i + 1] == "*":
            j = text.find("*/", i + 2)
            i = n if j < 0 else j + 2
        elif c in "\"'":
            j = i + 1
            while j < n and text[j] != c:
                j += 2 if text[j] == "\\" else 1
            out.append(text[i:j + 1])
            i = j + 1
        elif c.isspace():
            if out and out[-1] != " ":
                out.append(" ")                     # token separation survives as one blank
            i += 1
        else:
            out.append(c)
            i += 1
    return "".join(out)


def source_hash(csrc: str) -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(csrc, name), encoding="utf-8") as f:
                h.update(name.encode() + b"\0" + strip_code(f.read()).encode() + b"\0")
    return h.hexdigest()[:16]


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    print(source_hash(sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "fake-video-detection-engine_b200", "csrc")))
