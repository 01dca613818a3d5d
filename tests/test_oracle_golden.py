"""Pin the Python restatement oracle against the reference's own test vectors (tests/golden/*.json,
extracted from the Go tests by tools/gen_golden.py; every table carries its reference file:line)."""
import pytest

from oracle import pyoracle as O

EV = {"Unknown": 0, "Info": 1, "Warning": 2, "Critical": 3, "Fatal": 4}
ACT = {"IgnoreNoActionRequired": 1, "RebootSystem": 2, "HardwareInspection": 3, "CheckUserAppAndGPU": 4}


def test_extract_xid(golden):                      # xid/kmsg_test.go:14-55
    for r in golden("xid_kmsg.json")["extract_xid"]["rows"]:
        assert O.extract_nvrm_xid_info(r["input"].encode())[0] == r["expected"], r["name"]


def test_extract_device(golden):                   # xid/kmsg_test.go:57-103
    for r in golden("xid_kmsg.json")["extract_device"]["rows"]:
        assert O.extract_nvrm_xid_info(r["input"].encode())[1] == r["expected"], r["name"]


def test_match(golden):                            # xid/kmsg_test.go:105-246
    g = golden("xid_kmsg.json")
    for r in g["match"]["rows"] + g["unknown_code"]["rows"]:
        m = O.xid_match(r["input"].encode())
        if r.get("expectNil"):
            assert m is None, r
        else:
            assert m is not None and m.xid == r["expectedXid"] and m.device == r["expectedDevice"], r["name"]
            assert m.detail is not None


def test_dmesg_fixture(golden):                    # xid/kmsg_test.go:248-287
    g = golden("xid_kmsg.json")["dmesg_xid_119"]
    hits = [O.xid_match(l.encode()) for l in g["lines"]]
    hits = [h for h in hits if h is not None]
    assert [(h.xid, h.device) for h in hits] == [(r["xid"], r["device"]) for r in g["rows"]]
    # the buffer-scan form gives the same answer
    sc = O.scan_lines("\n".join(g["lines"]).encode())
    assert [(h["code"], h["device"]) for h in sc if h["kind"] == 1] == [(119, "PCI:0000:9b:00")] * 5


def test_normalize_bdf(golden):                    # xid/kmsg_test.go:312-352
    for r in golden("xid_kmsg.json")["normalize_bdf"]["rows"]:
        assert O.normalize_pci_bdf(r["input"]) == r["expected"], r["name"]


def test_extended(golden):                         # xid/kmsg_extended_test.go:15-326
    for r in golden("xid_kmsg.json")["extended"]["rows"]:
        info = O.extract_nvrm_xid_info_extended(r["logLine"].encode())
        if not r["shouldMatch"]:
            assert info is None, r["name"]
            continue
        assert info is not None, r["name"]
        assert info.xid == r["expectedXid"] and info.device == r["expectedDeviceUUID"]
        assert info.sub_code == r["expectedSubCode"] and info.unit == r["expectedSubCodeName"]
        assert info.severity == r["expectedSeverity"] and info.link == r["expectedLink"]
        assert info.intrinfo == r["expectedIntrinfo"] and info.error_status == r["expectedErrorStatus"]


def test_subcode_and_short(golden):                # xid/kmsg_extended_test.go:328-354
    g = golden("xid_kmsg.json")
    for r in g["subcode"]["rows"]:
        assert (r["intrinfo"] >> 20) & 0x3F == r["expectedCode"]
    for r in g["short_match"]["rows"]:
        assert O.extract_nvrm_xid_info_extended(r["line"].encode()) is None


def test_detail_with_subcode(golden):              # xid/kmsg_extended_test.go:356-464
    for r in golden("xid_kmsg.json")["detail_with_subcode"]["rows"]:
        d = O.get_detail_with_sub_code(r["xid"], r["subCode"])
        assert (d is not None) == r["expectedFound"]
        if d is None:
            continue
        if r["expectedEventTypeFatal"]:
            assert d.event_type == O.EV_FATAL, r["name"]
        assert d.code == r["xid"]


def test_match_nvlink_examples(golden):            # xid/kmsg_extended_test.go:466-524
    for r in golden("xid_kmsg.json")["match_nvlink_examples"]["rows"]:
        m = O.xid_match(r["logLine"].encode())
        assert m is not None, r["name"]
        assert m.detail.sub_code == r["expectedSub"], r["name"]
        assert m.detail.sub_code_description == r["expectedDesc"], r["name"]
        assert m.detail.event_type == EV[r["expectedEvent"]], r["name"]
        if r["expectedAction"] is not None:
            assert m.detail.actions is not None
            for a in r["expectedAction"]:
                assert ACT[a] in m.detail.actions, r["name"]


def test_nvlink_log_coverage(golden):              # xid/nvlink_logs_test.go:15-103
    for r in golden("xid_kmsg.json")["nvlink_log_coverage"]["rows"]:
        m = O.xid_match(r["line"].encode())
        assert m is not None and m.detail.event_type == EV[r["expectedEvent"]], r["name"]
        assert O.MNEMONIC[m.xid] != ""


def test_status_specific(golden):                  # xid/xid_test.go:13-55
    for r in golden("xid_kmsg.json")["status_specific"]["rows"]:
        info = O.ExtractedInfo("", r["xid"], "", "", r["unit"], r["severity"], "", "", 0, r["intrinfo"], r["error_status"], [], 0)
        d = O.detail_from_nvlink_info(info)
        assert d is not None and d.event_type == EV[r["event"]]
        for a in r.get("actions_contain", []):
            assert ACT[a] in d.actions


def test_inject_messages(golden):                  # xid/kmsg_test.go:432-515 + kmsg.go:278-311
    g = golden("xid_kmsg.json")["inject_messages"]
    for code, m in g["known"].items():
        assert O.extract_nvrm_xid_info(m["message"].encode())[0] == int(code)
    for code in (1, 25, 42, 100, 999):
        assert O.extract_nvrm_xid_info((g["template"] % code).encode())[0] == code


def test_sxid(golden):                             # sxid/kmsg_test.go:7-170
    g = golden("sxid_kmsg.json")
    for r in g["extract_sxid"]["rows"]:
        assert O.extract_sxid(r["input"].encode()) == r["expected"], r["name"]
    for r in g["extract_device"]["rows"]:
        assert O.extract_sxid_device(r["input"].encode()) == r["expected"], r["name"]
    for r in g["match"]["rows"]:
        m = O.sxid_match(r["input"].encode())
        if r.get("expectNil"):
            assert m is None, r["name"]
        else:
            assert m["sxid"] == r["expectedSXid"] and m["device"] == r["expectedDevice"], r["name"]


def test_catalog_checksums():                      # SURVEY.md A.3 (derived from xid.go:122-2952, sxid.go:94-2380)
    fatal = {9, 12, 16, 18, 19, 26, 27, 28, 29, 30, 32, 33, 34, 35, 36, 38, 42, 44, 46, 47, 48, 59, 60, 61, 62, 63, 64, 65,
             68, 69, 74, 78, 79, 81, 94, 95, 110, 119, 120, 121, 123, 140, 143}
    assert {c for c, d in O.XID_DETAILS.items() if d.event_type == O.EV_FATAL} == fatal
    assert {c for c, d in O.XID_DETAILS.items() if d.event_type == O.EV_INFO} == {152, 153, 157, 161, 162, 164, 165}
    assert sorted(O.XID_DETAILS) == [c for c in range(1, 174) if c != 133]
    assert O.XID_DETAILS[31].actions == [O.ACT_CHECK_APP, O.ACT_HW_INSPECTION]
    assert O.XID_DETAILS[94].actions == [O.ACT_REBOOT, O.ACT_IGNORE]
    assert len(O.SXID_DETAILS) == 93 and len(O.NVLINK_RULES) == 94


def test_parse_kmsg_line(golden):                  # pkg/kmsg/watcher_test.go (Test_parseLineComprehensive)
    import re
    rows = golden("pkg_kmsg.json")["Test_parseLineComprehensive"]["rows"]
    assert rows
    for r in rows:
        if r.get("expectError"):
            with pytest.raises(ValueError):
                O.parse_kmsg_line(0, r["input"])
            continue
        prio, seq, ts, msg = O.parse_kmsg_line(0, r["input"])
        e = r["expected"]
        assert (prio, seq, msg) == (e["Priority"], e["SequenceNumber"], e["Message"]), r["name"]
        m = re.search(r"bootTime\.Add\((\d+)\*time\.Microsecond", e["Timestamp"]["$call"])
        assert ts == (int(m.group(1)) if m else 0), r["name"]          # metav1.NewTime(bootTime) == +0us


def test_ext_matchers(golden):                     # nccl/kmsg_matcher_test.go:5,67 ; peermem/kmsg_matcher_test.go:5,62
    G = golden("ext_kmsg.json")
    for r in G["nccl_has"]["rows"]:
        assert (3 in O.ext_match(r["line"].encode())) == r["want"], r
    for r in G["nccl_match"]["rows"]:
        assert (3 in O.ext_match(r["line"].encode())) == bool(r["wantName"]), r
    for r in G["peermem_has"]["rows"]:
        assert (4 in O.ext_match(r["line"].encode())) == r["want"], r
    for r in G["peermem_match"]["rows"]:
        assert (4 in O.ext_match(r["line"].encode())) == bool(r["wantName"]), r


def test_ext2_matchers(golden):                    # infiniband / cpu / os / disk kmsg_matcher_test.go tables
    G = golden("ext2_kmsg.json")
    kinds = {}
    for comp in ("infiniband", "cpu", "os", "disk"):
        for pat in G[comp + ".patterns"]["patterns"]:
            k = [k for k, c, e, *_ in O.EXT_PATTERNS if c == comp and e == pat["event"]]
            assert len(k) == 1, pat
            assert O.EXT_PATTERNS[k[0] - 3][4].decode() == pat["regex"] and O.EXT_PATTERNS[k[0] - 3][3] == pat["message"], pat
            kinds[comp + "." + pat["key"]] = k[0]
    n = 0
    for key, kind in kinds.items():
        for r in G[key]["rows"]:
            want = r.get("want", r.get("expectedMatch"))
            assert (kind in O.ext_match(r["line"].encode())) == want, (key, r)
            if want and "expectedProcess" in r:
                assert O.ext_capture(kind, r["line"].encode()).decode() == r["expectedProcess"], r
            n += 1
    for comp in ("infiniband", "cpu", "os", "disk"):
        for r in G[comp + ".match"]["rows"]:
            assert O.component_match(comp, r["line"].encode()) == (r["wantEvent"], r["wantMessage"]), (comp, r)
            n += 1
    for r in G["infiniband.access_reg_message"]["rows"]:
        assert O.ext_message(7, r["line"].encode()) == r["want"], r
        n += 1
    assert n > 140


def test_stateful_matchers(golden):                # os kernel-panic assembly + memory OOM parser (kmsg_matcher_test.go tables)
    G = golden("ext3_kmsg.json")
    for r in G["os.panic_start"]["rows"]:
        assert bool(O.EXT_RE[19].search(r["line"].encode())) == r["want"], r
    for r in G["os.panic_cpu_pid"]["rows"]:
        m = O.EXT_RE[20].search(r["line"].encode())
        ok = bool(m) and O.go_atoi(m.group(1)) is not None and O.go_atoi(m.group(2)) is not None
        assert ok == r["wantFound"], r
        if ok:
            assert (int(m.group(1)), int(m.group(2)), m.group(3).decode()) == (r["wantCPU"], r["wantPID"], r["wantProcess"]), r
    for r in G["os.panic_detection"]["rows"]:
        pm, got = O.KernelPanicMatcher(), ("", "")
        for l in r["logLines"]:
            ev = pm.feed(l.encode())
            if ev[0]:
                got = ev
                break
        assert got == (r["wantEventName"], r["wantMessage"]), (r["name"], got)
    for r in G["os.panic_stateful"]["rows"]:
        pm = O.KernelPanicMatcher()
        for line, ev, msg in r["scenario"]:
            assert pm.feed(line.encode()) == (ev, msg), (r["name"], line)
    pm = O.KernelPanicMatcher()
    assert pm.feed(b"Kernel panic - not syncing: hung_task: blocked tasks") == ("", "")
    for _ in range(G["os.max_lines"]["max_lines_after_start"] - 1):
        assert pm.feed(b"line without cpu pid") == ("", "")
    assert pm.feed(b"10th line without cpu pid") == (G["os.max_lines"]["event"], G["os.max_lines"]["fallback_message"])
    for key in ("memory.match_func", "memory.stream"):
        for r in G[key]["rows"]:
            om = O.OOMMatcher()
            got = [ev for ev in (om.feed(l.encode()) for l in r["messages"]) if ev[0] or ev[1]]
            assert got == [(e["eventName"], e["message"]) for e in r["expectedEvents"]], (r["name"], got)
    for r in G["memory.oom_start"]["rows"]:
        assert bool(O.EXT_RE[21].search(r["line"].encode())) == r["expected"], r
    for r in G["memory.container_name"]["rows"]:
        om = O.OOMMatcher()
        om.feed(b"x invoked oom-killer:")
        om.feed(r["line"].encode())
        m = O.EXT_RE[22].search(r["line"].encode())
        assert bool(m) == (r["expectedFound"] or r["expectedError"]), r
        if r["expectedError"]:
            assert O.go_atoi(m.group(8)) is None and om.cur is None, r
        elif "expectedContainer" in r and om.cur is not None:
            assert (om.cur["container"], om.cur["victim"]) == (r["expectedContainer"], r["expectedVictim"]), r
        elif r["expectedFound"]:
            assert (m.group(6).decode(), m.group(5).decode(), m.group(1).decode(), int(m.group(8)), m.group(7).decode()) == \
                   (r["expectedContainer"], r["expectedVictim"], r["expectedConstraint"], r["expectedPid"], r["expectedProcess"]), r
    for r in G["memory.process_pid"]["rows"]:
        m = O.EXT_RE[24].search(r["line"].encode())
        assert bool(m) == (r["expectedFound"] or r["expectedError"]), r
        if r["expectedError"]:
            assert O.go_atoi(m.group(1)) is None, r
        elif r["expectedFound"]:
            assert (int(m.group(1)), m.group(2).decode()) == (r["expectedPid"], r["expectedProcess"]), r
    for r in G["memory.summary"]["rows"]:
        inst = r["instance"]
        if inst is None:
            continue                               # nil receiver: "" (Go-only case)
        inst = inst if isinstance(inst, dict) else {}
        o = {"pid": inst.get("Pid", 0), "process": inst.get("ProcessName", ""), "container": inst.get("ContainerName", ""),
             "victim": inst.get("VictimContainerName", ""), "constraint": inst.get("Constraint", "")}
        assert O.OOMMatcher.summary(o) == r["expected"], r


def test_ib_drop_flap_scans(golden):               # infiniband/store/scan_drops_test.go:15, scan_flaps_test.go:15
    from oracle import ib_scans as IB
    G = golden("ib_scans.json")
    base = 1_700_000_000
    def series(r):
        return [(base + int(x["t"]), x["state"] != "active", x["total_link_downed"]) for x in r["snapshots"]]
    for r in G["drops"]["rows"]:
        got = IB.find_drops(series(r), G["drops"]["threshold_s"])
        assert (1 if got else 0) == r["expected"], r["name"]
    for r in G["flaps"]["rows"]:
        got = IB.find_flaps(series(r), G["flaps"]["down_interval_threshold_s"], G["flaps"]["flap_back_to_active_threshold"])
        assert (1 if got else 0) == r["expected"], r["name"]
    assert sum(r["expected"] for r in G["drops"]["rows"]) >= 3 and sum(r["expected"] for r in G["flaps"]["rows"]) >= 3
    for r in G["edge"]["rows"]:                    # the t.Run sub-tests: per-call thresholds, Len and index asserts
        if r["kind"] == "drops":
            got = IB.find_drops(series(r), int(r["args"][0]))
        else:
            got = IB.find_flaps(series(r), int(r["args"][0]), int(r["args"][1]))
        assert (1 if got else 0) == r["expected"], r["name"]
        if "expected_index" in r:
            assert got["index"] == r["expected_index"], r["name"]
    import gpud_b200 as g                          # host-side reason formatting (no GPU needed)
    for t in (0, 1, 86399, 86400, 951782400, 1_700_000_000, 1709251199, 1709251200, 4102444800):
        assert g.ib_reason("mlx5_0", 1, t, False) == IB.drop_reason("mlx5_0", 1, t)
        assert g.ib_reason("mlx5_10", 2, t, True) == IB.flap_reason("mlx5_10", 2, t)
