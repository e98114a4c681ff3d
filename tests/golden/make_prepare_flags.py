"""Extracts the command-line contract of the reference's 0-prepare-coex.py (flags, destinations, defaults, types,
choices; lines 23-73) by parsing its source text -- the script itself is Python 2 and cannot be imported -- and
commits it as tests/golden/reference_prepare_flags.json for tests/test_prepare.py.  Run in the build container:
    python tests/golden/make_prepare_flags.py
"""
import ast
import json
import os
import re

REF = "/root/reference/0-prepare-coex.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_prepare_flags.json")


def main():
    src = open(REF).read()
    flags = []
    # every live (not commented) parser.add_argument(...) call, possibly spanning lines
    for m in re.finditer(r"^parser\.add_argument\((.*?)\)\s*$", src, re.S | re.M):
        call = ast.parse("f(" + m.group(1) + ")").body[0].value
        names = [a.value for a in call.args]
        kw = {}
        for k in call.keywords:
            if k.arg in ("default", "choices", "action"):
                try:
                    kw[k.arg] = ast.literal_eval(k.value)
                except Exception:
                    kw[k.arg] = "<expr>"              # e.g. default=version, default=os.path.join(...)
            elif k.arg == "type":
                kw["type"] = k.value.id
        flags.append({"names": names, **kw})
    # keys 0-prepare adds to settings.txt besides vars(args) (lines 443-450)
    extra = re.findall(r'settingsdict\["(\w+)"\]', src)
    with open(OUT, "w") as o:
        json.dump({"flags": flags, "settings_extra_keys": extra}, o, indent=1, sort_keys=True)
    print(len(flags), "flags;", extra)


if __name__ == "__main__":
    main()
