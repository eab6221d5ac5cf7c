#!/usr/bin/env python3
"""Oracle build helper (test infrastructure, never on the product path).

Expands the reference's `.t` templates straight from /root/reference into
oracle/_ref/gen/ (git-ignored).  The reference's own generators
(gnuradio-core/src/lib/filter/generate_gr_fir_XXX.py:32-65,
generate_gr_fir_filter_XXX.py:37-42, generate_gr_freq_xlating_fir_filter_XXX.py:38-42,
gnuradio-core/src/python/build_utils.py:179-192) are Python 2 and do not run here; this
applies the same @KEY@ substitutions.  No reference source is stored in this repo.
"""
import os
import re
import sys

TYPE = {"s": "short", "i": "int", "f": "float", "c": "gr_complex"}


def std_dict(name, code3):
    return {
        "NAME": name,
        "GUARD_NAME": "INCLUDED_%s_H" % name.upper(),
        "BASE_NAME": re.sub("^gr_", "", name),
        "SPTR_NAME": "%s_sptr" % name,
        "WARNING": "machine generated from the reference template",
        "COPYRIGHT": "",
        "TYPE": TYPE[code3[0]],
        "I_TYPE": TYPE[code3[0]],
        "O_TYPE": TYPE[code3[1]],
        "TAP_TYPE": TYPE[code3[2]],
    }


def fir_dict(name, code3):
    d = std_dict(name, code3)
    d["FIR_TYPE"] = "gr_fir_" + code3
    d["INPUT_CAST"] = "(float)" if (code3[0] == "s" and code3[1] == "c") else ""
    acc = "c" if "c" in code3 else ("f" if "f" in code3 else "i")
    d["ACC_TYPE"] = TYPE[acc]
    d["N_UNROLL"] = "2" if acc == "c" else "4"
    d["VRCOMPLEX_INCLUDE"] = "#include <gr_types.h>" if acc == "c" else ""
    return d


def expand(src, dst, d):
    text = open(src).read()
    text = re.sub(r"@([A-Z0-9_]+)@", lambda m: d[m.group(1)], text)
    with open(dst, "w") as f:
        f.write(text)


def main(ref, out):
    filt = os.path.join(ref, "gnuradio-core/src/lib/filter")
    os.makedirs(out, exist_ok=True)
    for code3 in ("ccf", "fff", "ccc"):
        for root in ("gr_fir_XXX", "gr_fir_XXX_generic"):
            name = root.replace("XXX", code3)
            d = fir_dict(name, code3)
            for ext in (".h", ".cc"):
                expand(os.path.join(filt, root + ext + ".t"), os.path.join(out, name + ext), d)
    for code3 in ("ccf", "fff"):
        name = "gr_fir_filter_" + code3
        d = std_dict(name, code3)
        d["FIR_TYPE"] = "gr_fir_" + code3
        for ext in (".h", ".cc"):
            expand(os.path.join(filt, "gr_fir_filter_XXX" + ext + ".t"), os.path.join(out, name + ext), d)
    name = "gr_freq_xlating_fir_filter_ccf"
    d = std_dict(name, "ccf")
    d["FIR_TYPE"] = "gr_fir_ccc"
    for ext in (".h", ".cc"):
        expand(os.path.join(filt, "gr_freq_xlating_fir_filter_XXX" + ext + ".t"), os.path.join(out, name + ext), d)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
