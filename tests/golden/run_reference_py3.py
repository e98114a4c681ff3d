"""Runs the REFERENCE'S OWN Python (`/root/reference/1-make-network.py` + `coexfunctions.py`) in this container.

TEST INFRASTRUCTURE -- used only by tests/golden/make_reference_golden.py (fixture generator, run where
/root/reference exists) to pin oracle/network_oracle.py against executed reference code.

The reference drivers are Python 2.7; this image has Python 3.12 only.  `coexfunctions.py` already
parses as Python 3 and is IMPORTED UNMODIFIED from /root/reference (matplotlib, which it imports but the path never
calls, is stubbed; `string.split` / `string.strip`, removed in Python 3, are put back).  `1-make-network.py` is
executed from its own text after a purely mechanical, line-local 2 -> 3 rewrite (nothing is copied into the repo):

    print statements            -> print(...)           (logical lines, continuation aware)
    d.iteritems()               -> d.items()
    xrange                      -> range
    lambda (k, v): (v, k)       -> lambda kv: (kv[1], kv[0])
    X.keys()                    -> list(X.keys())       (Python 2 `keys()` is an indexable snapshot)
    import ConfigParser         -> import configparser as ConfigParser

No arithmetic, control flow or data structure is touched.  What Python 3 changes on its own and what that means:
  * dict order is insertion order (Python 2: hash order).  The reference's results depend on dict order only through
    quirk Q6 (group order of the second winner pass, first-maximum ties).  make_reference_golden.py therefore compares
    order-independent outputs exactly and the order-dependent ones under the order this run used.
  * the built-in `sum` of Python >= 3.12 is a compensated (Neumaier) float sum; Python 2.7's is the plain
    left-to-right loop.  The run gets a `sum` with the 2.7 behaviour (`py27_sum`), otherwise the densities would differ
    from a real Python-2 run in the 14th digit.
  * scipy 1.x instead of 0.19: `spearmanr` may differ in the last bit -- the *faithful* oracle mode calls the same
    scipy, so both sides see the same numbers.
The C library loaded by the script is oracle/_ref/coexpression_v2.so = the unmodified reference C file.
"""
import os
import re
import runpy
import shutil
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _balanced(s):
    depth = 0
    quote = None
    i = 0
    while i < len(s):
        ch = s[i]
        if quote:
            if ch == "\\":
                i += 1
            elif ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
        elif ch == "#":
            break
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        i += 1
    return depth


def _comment_start(s):
    """index of a trailing # comment outside string literals (len(s) if none)"""
    quote = None
    i = 0
    while i < len(s):
        ch = s[i]
        if quote:
            if ch == "\\":
                i += 1
            elif ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
        elif ch == "#":
            return i
        i += 1
    return len(s)


_PRINT = re.compile(r"^(?P<head>\s*(?:(?:if|elif)\s[^:]*:\s*|else:\s*)?)print(?P<sp>\s+)(?P<rest>\S.*)$")


def convert_py2(src):
    out = []
    lines = src.split("\n")
    i = 0
    while i < len(lines):
        line = lines[i]
        m = _PRINT.match(line)
        if m and not m.group("rest").startswith("("):
            rest = m.group("rest")
            while _balanced(rest) > 0 or rest.rstrip().endswith("\\"):      # continuation lines of the statement
                i += 1
                rest = rest.rstrip().rstrip("\\") + " " + lines[i].strip()
            cut = _comment_start(rest)
            comment = "  " + rest[cut:] if cut < len(rest) else ""
            line = "%sprint(%s)%s" % (m.group("head"), rest[:cut].rstrip(), comment)
        out.append(line)
        i += 1
    s = "\n".join(out)
    s = s.replace(".iteritems()", ".items()")
    s = re.sub(r"\bxrange\(", "range(", s)
    s = s.replace("lambda (k, v): (v, k)", "lambda kv: (kv[1], kv[0])")
    s = re.sub(r"\b([A-Za-z_][A-Za-z_0-9]*)\.keys\(\)", r"list(\1.keys())", s)
    s = s.replace("import ConfigParser", "import configparser as ConfigParser")
    return s


def py27_sum(values, start=0):
    """Python 2.7 `sum`: start + v0 + v1 + ... strictly left to right, one rounding per addition."""
    acc = start
    for v in values:
        acc = acc + v
    return acc


def install_py2_shims():
    import string
    if not hasattr(string, "split"):
        string.split = lambda s, sep=None, maxsplit=-1: s.split(sep, maxsplit)
        string.strip = lambda s, chars=None: s.strip(chars)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if REF not in sys.path:
        sys.path.insert(0, REF)


def import_reference_coexfunctions():
    """The reference's coexfunctions module, unmodified."""
    install_py2_shims()
    import importlib
    mod = importlib.import_module("coexfunctions")
    assert os.path.dirname(os.path.abspath(mod.__file__)) == REF, mod.__file__
    mod.sum = py27_sum
    return mod


def run_make_network(appdir, working_dir, mapfile):
    """Executes the converted 1-make-network.py as __main__ with `-mf mapfile -wd working_dir`.  appdir receives the
    converted script and a link to the reference C build (the script loads `<its own dir>/coexpression_v2.so`)."""
    install_py2_shims()
    os.makedirs(appdir, exist_ok=True)
    so = os.path.join(ROOT, "oracle", "_ref", "coexpression_v2.so")
    assert os.path.exists(so), "make -C oracle first (builds the unmodified reference C file)"
    dst = os.path.join(appdir, "coexpression_v2.so")
    if not os.path.exists(dst):
        shutil.copy(so, dst)
    script = os.path.join(appdir, "1-make-network.py")
    with open(os.path.join(REF, "1-make-network.py")) as f:
        converted = convert_py2(f.read())
    with open(script, "w") as o:
        o.write(converted)
    argv = sys.argv
    sys.argv = [script, "-mf", mapfile, "-wd", working_dir]
    try:
        import_reference_coexfunctions()
        return runpy.run_path(script, init_globals={"sum": py27_sum}, run_name="__main__")
    finally:
        sys.argv = argv


def run_collate(appdir, working_dir):
    """Executes the converted 2-collate-results.py as __main__ with `-wd working_dir`."""
    install_py2_shims()
    os.makedirs(appdir, exist_ok=True)
    script = os.path.join(appdir, "2-collate-results.py")
    with open(os.path.join(REF, "2-collate-results.py")) as f:
        converted = convert_py2(f.read())
    with open(script, "w") as o:
        o.write(converted)
    argv = sys.argv
    sys.argv = [script, "-wd", working_dir]
    try:
        import_reference_coexfunctions()
        return runpy.run_path(script, init_globals={"sum": py27_sum}, run_name="__main__")
    finally:
        sys.argv = argv


if __name__ == "__main__":
    if sys.argv[1] == "collate":
        run_collate(sys.argv[2], sys.argv[3])
    else:
        run_make_network(sys.argv[1], sys.argv[2], sys.argv[3])
