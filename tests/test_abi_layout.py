"""The C header's structure layouts (offsetof, compiled with gcc) against the ctypes mirror, the oracle's mirror and
the offset table in INTEGRATION.md section 2 — what a `foreign import` binding pokes at (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

from massiv_b200 import _abi
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_layout(tmp_path):
    exe = os.path.join(str(tmp_path), "abi_layout")
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_layout.c"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return {ln.split()[0]: (int(ln.split()[1]), int(ln.split()[2])) for ln in out.splitlines()}


def test_ctypes_and_oracle_mirrors_match_the_header(tmp_path):
    lay = c_layout(tmp_path)
    for cname, mirrors in (("mb_stencil_desc", (_abi.StencilDesc, O._Desc)), ("mb_term", (_abi.Term, O._Term)),
                           ("mb_xins", (_abi.XIns, O._XIns)), ("mb_post_op", (_abi.PostOp,)), ("mb_shard_plan", (_abi.ShardPlan,))):
        for m in mirrors:
            assert C.sizeof(m) == lay[cname + ".sizeof"][0], (cname, m)
            for fname, _ in m._fields_:
                if fname == "reserved":
                    continue
                off, size = lay[f"{cname}.{fname}"]
                fld = getattr(m, fname)
                assert (fld.offset, fld.size) == (off, size), (cname, fname)


def test_integration_md_offset_table_matches_the_header(tmp_path):
    lay = c_layout(tmp_path)
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    rows = re.findall(r"^\| `(mb_\w+)\.(\w+)` \| (\d+) \| (\d+) \|", text, flags=re.M)
    assert len(rows) >= 30, "INTEGRATION.md section 2 lost its offset table"
    seen = set()
    for st, fld, off, size in rows:
        assert lay[f"{st}.{fld}"] == (int(off), int(size)), (st, fld)
        seen.add(f"{st}.{fld}")
    # every field of the descriptor and its sub-structures is documented
    for key in lay:
        if key.startswith(("mb_stencil_desc.", "mb_term.", "mb_xins.", "mb_post_op.")) and not key.endswith(".sizeof"):
            assert key in seen, key
    for st in ("mb_stencil_desc", "mb_term", "mb_xins", "mb_post_op"):
        m = re.search(rf"`{st}` is (\d+) bytes", text)
        assert m and int(m.group(1)) == lay[st + ".sizeof"][0], st
