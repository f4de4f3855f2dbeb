"""Regenerates the ctypes structure stub of INTEGRATION.md section 2 from mdn_sfm_b200/_cabi.py (the binding the product
uses), between the `<!-- stub:begin -->` / `<!-- stub:end -->` markers.  tests/test_cabi_exports.py executes the stub and
holds its sizeof / offsetof to the C compiler's view of include/mdn_loss.h."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdn_sfm_b200 import _cabi  # noqa: E402


def ctype_name(t):
    if hasattr(t, "_length_"):      # arrays
        base = t._type_
        return "%s * %d" % ("MdnScale" if base is _cabi.MdnScale else ctype_name(base), t._length_)
    return {C.c_int32: "C.c_int32", C.c_float: "C.c_float", C.c_double: "C.c_double", C.c_void_p: "C.c_void_p"}[t]


def struct_src(cls, comment):
    lines, cur = [], "    _fields_ = ["
    for name, t in cls._fields_:
        item = '("%s", %s), ' % (name, ctype_name(t))
        if len(cur) + len(item) > 118:
            lines.append(cur.rstrip())
            cur = "                "
        cur += item
    lines.append(cur.rstrip(", ") + "]")
    return "class %s(C.Structure):            # mirrors `struct %s`, include/mdn_loss.h%s\n%s\n" % (cls.__name__, cls.__name__, comment, "\n".join(lines))


def stub():
    return ("```python\nimport ctypes as C\n\n" + struct_src(_cabi.MdnScale, "") + "\n"
            + struct_src(_cabi.MdnLossDesc, " (ABI %d)" % _cabi.ABI_VERSION) + "```\n")


def main():
    path = os.path.join(ROOT, "INTEGRATION.md")
    s = open(path).read()
    a, b = s.index("<!-- stub:begin -->"), s.index("<!-- stub:end -->")
    s = s[:a] + "<!-- stub:begin -->\n" + stub() + s[b:]
    open(path, "w").write(s)
    print("INTEGRATION.md stub regenerated for ABI", _cabi.ABI_VERSION)


if __name__ == "__main__":
    main()
