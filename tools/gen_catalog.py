#!/usr/bin/env python3
"""Mechanically derive the Xid / SXid / NVLink-rule catalog tables from the reference tree.

Reads (never copies source code from):
  components/accelerator/nvidia/xid/xid.go:122-2952            `details` map
  components/accelerator/nvidia/xid/catalog_generated.go:7-181  `catalogEntries`
  components/accelerator/nvidia/xid/catalog_generated.go:183-277 `nvlinkRules`
  components/accelerator/nvidia/sxid/sxid.go:94-2380            `details` map

Writes DATA ONLY:
  oracle/catalog.json               (consumed by oracle/pyoracle.py and the tests)
  gpud_b200/csrc/catalog_data.inc   (C initialisers consumed by oracle/oracle.c and the product library)

Run in the build container (needs /root/reference); the outputs are committed because the GPU box
has no reference tree.  Checksums asserted at the bottom come from SURVEY.md appendix A.3.
"""
import json
import os
import re
import sys

REF = os.environ.get("GPUD_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NV = os.path.join(REF, "components/accelerator/nvidia")

EVENT = {"Unknown": 0, "Info": 1, "Warning": 2, "Critical": 3, "Fatal": 4}
ACTION = {  # api/v1/types.go:183-203 ; numeric ids are ours, names are the wire strings
    "IgnoreNoActionRequired": 1,
    "RebootSystem": 2,
    "HardwareInspection": 3,
    "CheckUserAppAndGPU": 4,
}


def go_unquote(lit: str) -> str:
    """Decode one Go string literal (interpreted "..." or raw `...`)."""
    lit = lit.strip()
    if lit.startswith("`"):
        return lit[1:-1]
    assert lit.startswith('"') and lit.endswith('"'), lit
    body = lit[1:-1]
    out, i = [], 0
    while i < len(body):
        c = body[i]
        if c != "\\":
            out.append(c)
            i += 1
            continue
        n = body[i + 1]
        simple = {"n": "\n", "t": "\t", '"': '"', "\\": "\\", "'": "'", "r": "\r"}
        if n in simple:
            out.append(simple[n])
            i += 2
        elif n == "x":
            out.append(chr(int(body[i + 2:i + 4], 16)))
            i += 4
        elif n == "u":
            out.append(chr(int(body[i + 2:i + 6], 16)))
            i += 6
        else:
            raise ValueError("escape \\%s" % n)
    return "".join(out)


STR = r'("(?:[^"\\]|\\.)*"|`[^`]*`)'


def split_map_entries(text: str, start_marker: str):
    """Yield (code, body) for `\\t<code>: {` ... `\\n\\t},` entries of a Go map literal."""
    start = text.index(start_marker)
    body = text[start:]
    # entry header may carry a trailing // comment (sxid 22012)
    heads = list(re.finditer(r"\n\t(\d+): \{[^\n]*\n", body))
    for idx, m in enumerate(heads):
        end = heads[idx + 1].start() if idx + 1 < len(heads) else len(body)
        chunk = body[m.end():end]
        # cut at the closing of this entry
        close = chunk.find("\n\t},")
        assert close >= 0
        yield int(m.group(1)), chunk[:close]


def strip_comments(chunk: str) -> str:
    out = []
    for line in chunk.split("\n"):
        # remove // comments that are outside string literals (entries never hold // inside "..." on the same
        # line as code except URLs in raw strings, handled by keeping raw strings intact below)
        s = line.lstrip()
        if s.startswith("//"):
            continue
        out.append(line)
    return "\n".join(out)


def parse_actions(chunk: str):
    m = re.search(r"RepairActions:\s*\[\]apiv1\.RepairActionType\{(.*?)\}", chunk, re.S)
    if not m:
        return []
    return [ACTION[a] for a in re.findall(r"apiv1\.RepairActionType(\w+)", m.group(1))]


def parse_xid_details():
    text = open(os.path.join(NV, "xid/xid.go")).read()
    out = {}
    for code, chunk in split_map_entries(text, "var details = map[int]Detail{"):
        c = strip_comments(chunk)
        assert int(re.search(r"Code:\s*(\d+)", c).group(1)) == code
        desc = go_unquote(re.search(r"Description:\s*" + STR, c).group(1))
        ev = re.search(r"EventType:\s*apiv1\.EventType(\w+)", c).group(1)
        out[code] = {"code": code, "description": desc, "event_type": EVENT[ev], "actions": parse_actions(c)}
    return out


def parse_struct_rows(text: str, var: str, fields):
    start = text.index("var %s = " % var)
    end = text.index("\n}\n", start)
    rows = []
    for line in text[start:end].split("\n"):
        line = line.strip()
        if not line.startswith("{"):
            continue
        row = {}
        for f, kind in fields:
            if kind == "s":
                m = re.search(r"\b%s: %s" % (f, STR), line)
                row[f] = go_unquote(m.group(1)) if m else ""
            elif kind == "i":
                m = re.search(r"\b%s: (0x[0-9a-fA-F]+|\d+)" % f, line)
                row[f] = int(m.group(1), 0) if m else 0
        rows.append(row)
    return rows


def parse_sxid_details():
    text = open(os.path.join(NV, "sxid/sxid.go")).read()
    out = {}
    for code, chunk in split_map_entries(text, "var details = map[int]Detail{"):
        c = strip_comments(chunk)
        m = re.search(r"SXid:\s*(\d+)", c)
        if m:
            assert int(m.group(1)) == code
        name = re.search(r"Name:\s*" + STR, c)
        ev = re.search(r"EventType:\s*apiv1\.EventType(\w+)", c).group(1)

        def flag(f):
            mm = re.search(r"%s:\s*(\w+(?:\.\w+)?)" % f, c)
            v = mm.group(1) if mm else "false"
            if v in ("true", "false"):
                return v == "true"
            # defaultPotentialFatalErr.{PotentialFatal=true,AlwaysFatal=false}; defaultAlwaysFatalErr.{true,true}
            tbl = {"defaultPotentialFatalErr.PotentialFatal": True, "defaultPotentialFatalErr.AlwaysFatal": False,
                   "defaultAlwaysFatalErr.PotentialFatal": True, "defaultAlwaysFatalErr.AlwaysFatal": True}
            return tbl[v]

        out[code] = {"sxid": code, "name": go_unquote(name.group(1)) if name else "",
                     "event_type": EVENT[ev], "actions": parse_actions(c),
                     "potential_fatal": flag("PotentialFatal"), "always_fatal": flag("AlwaysFatal")}
    return out


def c_str(s: str) -> str:
    out = []
    for ch in s.encode("utf-8"):
        if ch == ord('"'):
            out.append('\\"')
        elif ch == ord("\\"):
            out.append("\\\\")
        elif ch == ord("\n"):
            out.append("\\n")
        elif 32 <= ch < 127:
            out.append(chr(ch))
        else:
            out.append('\\%03o' % ch)
    return '"' + "".join(out) + '"'


def main():
    xid = parse_xid_details()
    gen = open(os.path.join(NV, "xid/catalog_generated.go")).read()
    entries = parse_struct_rows(gen, "catalogEntries", [("Code", "i"), ("Mnemonic", "s"), ("Description", "s"),
                                                        ("ImmediateResolution", "s"), ("InvestigatoryResolution", "s")])
    rules = parse_struct_rows(gen, "nvlinkRules", [("Xid", "i"), ("Unit", "s"), ("IntrinfoPatternV1", "s"),
                                                   ("IntrinfoPatternV2", "s"), ("ErrorStatus", "i"), ("Resolution", "s"),
                                                   ("Investigatory", "s"), ("Severity", "s")])
    sxid = parse_sxid_details()

    # ---- checksums from SURVEY.md A.3 ----
    assert sorted(xid) == [c for c in range(1, 174) if c != 133], len(xid)
    hist = {}
    for d in xid.values():
        hist[d["event_type"]] = hist.get(d["event_type"], 0) + 1
    assert hist == {EVENT["Fatal"]: 43, EVENT["Warning"]: 122, EVENT["Info"]: 7}, hist
    assert len(entries) == 172 and len(rules) == 94, (len(entries), len(rules))
    assert len(sxid) == 93
    sh = {}
    for d in sxid.values():
        sh[d["event_type"]] = sh.get(d["event_type"], 0) + 1
    assert sh == {EVENT["Fatal"]: 64, EVENT["Warning"]: 29}, sh

    cat = {"event_type_ids": EVENT, "action_ids": ACTION,
           "xid": [xid[k] for k in sorted(xid)],
           "catalog_entries": entries, "nvlink_rules": rules,
           "sxid": [sxid[k] for k in sorted(sxid)]}
    with open(os.path.join(ROOT, "oracle/catalog.json"), "w") as f:
        json.dump(cat, f, indent=1, sort_keys=True)
        f.write("\n")

    mn = {e["Code"]: e["Mnemonic"] for e in entries}
    L = ["/* GENERATED by tools/gen_catalog.py from the reference catalog tables - data only, do not edit.",
         " * xid:  components/accelerator/nvidia/xid/xid.go:122-2952, xid/catalog_generated.go:7-277",
         " * sxid: components/accelerator/nvidia/sxid/sxid.go:94-2380",
         " * event ids: 0 Unknown 1 Info 2 Warning 3 Critical 4 Fatal (api/v1/types.go:222-244)",
         " * action ids: 1 IGNORE_NO_ACTION_REQUIRED 2 REBOOT_SYSTEM 3 HARDWARE_INSPECTION 4 CHECK_USER_APP_AND_GPU */",
         "#define GPUD_CAT_N_XID %d" % len(xid), "#define GPUD_CAT_N_RULES %d" % len(rules),
         "#define GPUD_CAT_N_SXID %d" % len(sxid),
         "/* {code, event, n_actions, {a0..a3}, description, mnemonic} */",
         "static const gpud_cat_xid_row GPUD_CAT_XID[GPUD_CAT_N_XID] = {"]
    for k in sorted(xid):
        d = xid[k]
        a = d["actions"] + [0] * (4 - len(d["actions"]))
        L.append("  {%d, %d, %d, {%d,%d,%d,%d}, %s, %s}," % (k, d["event_type"], len(d["actions"]), a[0], a[1], a[2], a[3],
                                                           c_str(d["description"]), c_str(mn.get(k, ""))))
    L.append("};")
    L.append("/* {xid, unit, patV1, patV2, error_status, resolution, investigatory, severity} */")
    L.append("static const gpud_cat_rule_row GPUD_CAT_RULES[GPUD_CAT_N_RULES] = {")
    for r in rules:
        L.append("  {%d, %s, %s, %s, 0x%08xu, %s, %s, %s}," % (
            r["Xid"], c_str(r["Unit"]), c_str(r["IntrinfoPatternV1"]), c_str(r["IntrinfoPatternV2"]), r["ErrorStatus"],
            c_str(r["Resolution"]), c_str(r["Investigatory"]), c_str(r["Severity"])))
    L.append("};")
    L.append("/* {sxid, event, n_actions, {a0..a3}, potential_fatal, always_fatal, name} */")
    L.append("static const gpud_cat_sxid_row GPUD_CAT_SXID[GPUD_CAT_N_SXID] = {")
    for k in sorted(sxid):
        d = sxid[k]
        a = d["actions"] + [0] * (4 - len(d["actions"]))
        L.append("  {%d, %d, %d, {%d,%d,%d,%d}, %d, %d, %s}," % (k, d["event_type"], len(d["actions"]), a[0], a[1], a[2], a[3],
                                                               int(d["potential_fatal"]), int(d["always_fatal"]), c_str(d["name"])))
    L.append("};")
    with open(os.path.join(ROOT, "gpud_b200/csrc/catalog_data.inc"), "w") as f:
        f.write("\n".join(L) + "\n")
    # the C oracle gets its OWN copy (oracle/ compiles nothing from gpud_b200/): same generator, separate file, own row typedefs
    with open(os.path.join(ROOT, "oracle/oracle_catalog_data.inc"), "w") as f:
        f.write("/* TEST INFRASTRUCTURE: the oracle's copy of the generated catalog rows (tools/gen_catalog.py). */\n")
        f.write("typedef struct { int code; int event; int n_actions; int actions[4]; const char* description; const char* mnemonic; } gpud_cat_xid_row;\n")
        f.write("typedef struct { int xid; const char* unit; const char* pat_v1; const char* pat_v2; unsigned int error_status; const char* resolution;\n"
                "                 const char* investigatory; const char* severity; } gpud_cat_rule_row;\n")
        f.write("typedef struct { int sxid; int event; int n_actions; int actions[4]; int potential_fatal; int always_fatal; const char* name; } gpud_cat_sxid_row;\n")
        f.write("\n".join(L) + "\n")
    print("xid %d, rules %d, sxid %d -> oracle/catalog.json, gpud_b200/csrc/catalog_data.inc, oracle/oracle_catalog_data.inc" % (len(xid), len(rules), len(sxid)))


if __name__ == "__main__":
    sys.exit(main())
