"""The environment switches README.md lists are the ones the code reads, and the other way round (no GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts), errors="replace") as f:
        return f.read()


def _sources():
    out = []
    for d, exts in (("ggmlsharp_b200/csrc", (".cu", ".cuh", ".h")), ("ggmlsharp_b200/host", (".cpp", ".h")), ("ggmlsharp_b200", (".py",)),
                    ("benchmarks", (".py", ".sh")), (".", ("bench.py", "__graft_entry__.py"))):
        for name in sorted(os.listdir(os.path.join(ROOT, d))):
            if name.endswith(exts):
                out.append(_read(d, name))
    return "\n".join(out)


def test_every_documented_switch_is_read_somewhere():
    documented = set(re.findall(r"GGB200_[A-Z0-9_]+", _read("README.md")))
    src = _sources()
    missing = sorted(v for v in documented if v not in src)
    assert not missing, "README.md documents switches nothing reads: %s" % missing


def test_every_switch_the_library_reads_is_documented():
    lib = "\n".join(_read("ggmlsharp_b200/csrc", n) for n in sorted(os.listdir(os.path.join(ROOT, "ggmlsharp_b200/csrc"))) if n.endswith((".cu", ".cuh", ".h")))
    lib += _read("ggmlsharp_b200/host", "ggml_host.cpp")
    read = set(re.findall(r'getenv\("(GGB200_[A-Z0-9_]+)"\)', lib))
    docs = _read("README.md") + _read("DESIGN.md") + _read("INTEGRATION.md") + _read("profiles", "README.md")
    missing = sorted(v for v in read if v not in docs)
    assert not missing, "switches read by the library but documented nowhere: %s" % missing
