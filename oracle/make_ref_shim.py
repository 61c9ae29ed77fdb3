#!/usr/bin/env python
"""Materialise a Python-3 importable view of the (Python-2) reference in a SCRATCH directory.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline leg may touch it.

The reference (`/root/reference/learning`, Python 2 + NumPy) cannot be imported
by this image's Python 3.12 / NumPy 2.3 (``print`` statements, ``reduce``,
``reversed(zip(..))``, ``numpy.Infinity``).  This script applies the purely
syntactic edits listed in SURVEY.md §8(c) -- no arithmetic line changes -- and
writes the result OUTSIDE the repository (default ``/tmp/opl_ref_shim``), so no
reference source is ever copied into this repo.  The shim is used here, in the
build container, by ``oracle/gen_golden.py`` to

  * check that the NumPy restatement in ``oracle/`` reproduces the reference, and
  * emit the golden vectors committed under ``tests/golden/``.

It cannot travel to the GPU box (``/root/reference`` does not exist there);
everything run there uses the committed golden vectors and the oracle port.
"""
import os
import re
import shutil
import sys

REFERENCE = os.environ.get("OPL_REFERENCE_DIR", "/root/reference")
DEFAULT_OUT = os.environ.get("OPL_REF_SHIM_DIR", "/tmp/opl_ref_shim")

_NEEDS_REDUCE = re.compile(r"\breduce\(")


def _py3(text):
    lines = []
    for line in text.split("\n"):
        m = re.match(r"^(\s*)print (.*)$", line)
        if m and not m.group(2).lstrip().startswith("("):
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        lines.append(line)
    text = "\n".join(lines)
    # multi-line print continuation in testing/helpers.py: "print 'a' % (\n b, c)"
    text = re.sub(r"print\(('Actual shape[^\n]*\(\n[^\n]*\))\)", r"print(\1", text)
    text = re.sub(r"reversed\(\s*zip\(", "reversed(list_zip(", text)
    text = text.replace("numpy.Infinity", "numpy.inf")
    text = text.replace(".iteritems()", ".items()").replace(".itervalues()", ".values()")
    text = text.replace("collections.Iterable", "collections.abc.Iterable")
    text = text.replace("time.clock()", "time.perf_counter()")
    header = []
    if _NEEDS_REDUCE.search(text):
        header.append("from functools import reduce")
    if "list_zip(" in text:
        header.append("def list_zip(*a):\n    return list(zip(*a))")
    if "collections.abc" in text:
        header.append("import collections.abc")
    if header:
        # insert after the module docstring / licence block: before first import
        idx = re.search(r"^(import |from )", text, flags=re.M)
        pos = idx.start() if idx else 0
        text = text[:pos] + "\n".join(header) + "\n" + text[pos:]
    return text


def build(out_dir=DEFAULT_OUT):
    src = os.path.join(REFERENCE, "learning")
    if not os.path.isdir(src):
        raise RuntimeError("reference not present at %s" % src)
    dst = os.path.join(out_dir, "learning")
    if os.path.isdir(out_dir):
        shutil.rmtree(out_dir)
    for root, dirs, files in os.walk(src):
        rel = os.path.relpath(root, src)
        if rel.split(os.sep)[0] == "tests":
            continue
        os.makedirs(os.path.join(dst, rel), exist_ok=True)
        for name in files:
            s = os.path.join(root, name)
            d = os.path.join(dst, rel, name)
            if name.endswith(".py"):
                with open(s) as fh:
                    body = fh.read()
                with open(d, "w") as fh:
                    fh.write(_py3(body))
            else:
                shutil.copyfile(s, d)
    return out_dir


if __name__ == "__main__":
    out = build(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_OUT)
    sys.path.insert(0, out)
    import learning  # noqa: F401  (import check)
    print("shimmed reference importable from", out)
