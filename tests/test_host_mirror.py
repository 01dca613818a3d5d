"""The C++ host-side mirror of the reference's operator interface (gpud_b200/csrc/host_component.cpp), exercised through its
flat test entry points, against the reference's vectors and the Python oracle.  CPU only (no compute calls)."""
import ctypes as C

import pytest

import gpud_b200 as g
from oracle import pyoracle as O

ACT = {"IGNORE_NO_ACTION_REQUIRED": 1, "REBOOT_SYSTEM": 2, "HARDWARE_INSPECTION": 3, "CHECK_USER_APP_AND_GPU": 4}
EV = {"Unknown": 0, "Info": 1, "Warning": 2, "Critical": 3, "Fatal": 4}
HEALTH = ["Healthy", "Degraded", "Unhealthy"]


class HEvent(C.Structure):
    _fields_ = [("kind", C.c_int32), ("event_type", C.c_int32), ("xid", C.c_uint64), ("n_actions", C.c_int32), ("actions", C.c_int32 * 4)]


def test_parse_kmsg_line_matches_reference_vectors(golden):          # pkg/kmsg/watcher_test.go (Test_parseLineComprehensive)
    L = g.lib()
    rows = golden("pkg_kmsg.json")["Test_parseLineComprehensive"]["rows"]
    for r in rows:
        prio, seq, usec = C.c_int32(), C.c_int64(), C.c_int64()
        msg = C.create_string_buffer(8192)
        rc = L.gpudh_parse_kmsg_line(r["input"].encode(), C.byref(prio), C.byref(seq), C.byref(usec), msg, 8192)
        if r.get("expectError"):
            assert rc != 0, r["name"]
            with pytest.raises(ValueError):
                O.parse_kmsg_line(0, r["input"])
            continue
        assert rc == 0, r["name"]
        want = O.parse_kmsg_line(0, r["input"])
        assert (prio.value, seq.value, usec.value, msg.value.decode()) == want, r["name"]
        assert (prio.value, seq.value, msg.value.decode()) == (r["expected"]["Priority"], r["expected"]["SequenceNumber"], r["expected"]["Message"])


def test_parse_fixture_records(golden):                               # pkg/kmsg/testdata/kmsg.1.log
    L = g.lib()
    n = 0
    for rec in golden("pkg_kmsg.json")["fixture:kmsg.1.log"]["records"]:
        if not rec:
            continue
        prio, seq, usec = C.c_int32(), C.c_int64(), C.c_int64()
        msg = C.create_string_buffer(8192)
        rc = L.gpudh_parse_kmsg_line(rec.encode("utf-8", "surrogateescape"), C.byref(prio), C.byref(seq), C.byref(usec), msg, 8192)
        try:
            want = O.parse_kmsg_line(0, rec)
        except ValueError:
            assert rc != 0
            continue
        assert rc == 0 and (prio.value, seq.value, usec.value) == want[:3]
        n += 1
    assert n >= 40


def test_dedup_key_and_counts():                                      # pkg/kmsg/deduper.go:63-125, deduper_test.go
    L = g.lib()
    out = C.create_string_buffer(512)
    for t, m in ((1700000123, "hello"), (59, "x"), (60, "x"), (1700000160, "NVRM: Xid (PCI:0000:05:00): 79, a")):
        L.gpudh_dedup_key(C.c_int64(t), m.encode(), out, 512)
        assert out.value.decode() == O.dedup_key(t, m)
    L.gpudh_deduper_new.restype = C.c_void_p
    d = C.c_void_p(L.gpudh_deduper_new(C.c_int64(900)))
    add = lambda now, t, m: L.gpudh_deduper_add(d, C.c_int64(now), C.c_int64(t), m.encode())
    assert add(1000, 1000, "a") == 1 and add(1001, 1010, "a") == 2          # same minute bucket: second occurrence is dropped (>1)
    assert add(1002, 1070, "a") == 1                                         # next minute: new key
    assert add(1003, 1000, "b") == 1
    assert add(1000 + 901, 1000, "a") == 1                                   # TTL (15 min) expired
    L.gpudh_deduper_free(d)


def _evolve(events, threshold=2):
    L = g.lib()
    arr = (HEvent * max(1, len(events)))()
    for i, e in enumerate(events):
        if e["k"] == "xid":
            arr[i].kind, arr[i].event_type, arr[i].xid = 0, EV[e["type"]], e["xid"]
            acts = e.get("actions")
            arr[i].n_actions = -1 if acts is None else len(acts)
            for j, a in enumerate(acts or []):
                arr[i].actions[j] = ACT[a]
        else:
            arr[i].kind = 1 if e["k"] == "reboot" else 2
    h, a, x = C.c_int32(), C.c_int32(), C.c_uint64()
    L.gpudh_evolve(arr, len(events), threshold, C.byref(h), C.byref(a), C.byref(x))
    return HEALTH[h.value], a.value, x.value


def test_evolve_healthy_state_reference_scenarios(golden):            # xid/health_state_test.go:78-246
    for r in golden("xid_health.json")["rows"]:
        health, action, _ = _evolve(r["events"])
        assert health == r["health"], r["name"]
        assert action == (ACT[r["action"]] if r["action"] else 0), r["name"]
        # the Python oracle agrees
        ev = [{"name": "error_xid", "type": e["type"], "xid": e["xid"], "actions": None if e.get("actions") is None else [ACT[a] for a in e["actions"]]}
              if e["k"] == "xid" else {"name": "reboot"} for e in r["events"]]
        o = O.evolve_healthy_state(ev, 2)
        assert o["health"] == r["health"] and (o["actions"][0] if o["actions"] else 0) == action, r["name"]


def test_component_name_matches_reference():
    L = g.lib()
    L.gpudh_xid_component_name.restype = C.c_char_p
    assert L.gpudh_xid_component_name() == b"accelerator-nvidia-error-xid"        # xid/component.go:35


def test_hw_slowdown_window_rule():               # hw-slowdown/component.go:352-407 (defaults 10 min / 0.6 per minute, :29-35)
    L = g.lib()
    now = 1_700_000_000
    cases = [([now - 60 * i for i in range(1, 7)], 600, 0.6, "Unhealthy"),          # 6 distinct minutes / 10 = 0.6 >= 0.6
             ([now - 60 * i for i in range(1, 6)], 600, 0.6, "Healthy"),            # 5 / 10 < 0.6
             ([now - 61, now - 62, now - 63, now - 119], 600, 0.6, "Healthy"),      # same minute counts once
             ([now - 700], 600, 0.6, "Healthy"),                                    # outside the window
             ([now - 5], 0, 0.6, "Healthy"),                                        # no evaluation window
             ([now - 30 * i for i in range(1, 21)], 600, 0.6, "Unhealthy")]
    for ev, win, thr, want in cases:
        arr = (C.c_int64 * max(1, len(ev)))(*ev)
        freq, distinct = C.c_double(), C.c_int32()
        h = L.gpudh_hw_slowdown(arr, len(ev), C.c_int64(now), C.c_int64(win), C.c_double(thr), C.byref(freq), C.byref(distinct))
        oh, ofreq, on = O.hw_slowdown_state(ev, now, win, thr)
        assert HEALTH[h] == want == oh and distinct.value == on and abs(freq.value - ofreq) < 1e-12
        out = C.create_string_buffer(400)
        insp = L.gpudh_hw_slowdown_reason(arr, len(ev), C.c_int64(now), C.c_int64(win), C.c_double(thr), out, 400)
        full = O.hw_slowdown_check(ev, now, win, thr)
        assert out.value.decode() == full[3] and bool(insp) == full[4] == (want == "Unhealthy")
    # the bucket read is strictly newer than (now - window) (pkg/eventstore/database.go:330 "timestamp > ?")
    assert O.hw_slowdown_check([now - 600], now, 600, 0.05)[3] == "no clock events found"
    assert O.hw_slowdown_check([now - 599], now, 600, 0.05)[0] == "Unhealthy"
    # TestHighFrequencySlowdownEvents (hw-slowdown/component_test.go:862-896): 10 events one per minute back from now, window 10 min, 0.6
    r = O.hw_slowdown_check([now - 60 * i for i in range(10)], now, 600, 0.6)
    assert r[0] == "Unhealthy" and "hw slowdown events frequency per minute" in r[3] and "exceeded threshold" in r[3] and r[3].endswith("for the last 10m0s")
    # negative window (TestComponentStatesEdgeCases :446): since lies in the future, nothing is read
    assert O.hw_slowdown_check([now - 300], now, -600, 0.6)[3] == "no clock events found"
    for sec, want in [(0, "0s"), (45, "45s"), (600, "10m0s"), (3600, "1h0m0s"), (5405, "1h30m5s"), (-600, "-10m0s"), (86400 * 3, "72h0m0s")]:
        assert O.go_duration_seconds(sec) == want
        L.gpudh_go_duration.restype = None
        L.gpudh_go_duration(C.c_int64(sec), out, 400)
        assert out.value.decode() == want


def test_temperature_rule():                      # temperature/component.go:206-248: strict '>' on limits, '<=' on the margin
    L = g.lib()
    T = lambda cur, gmax, hbm, mmax, margin, mthr, hs=1, sl=95, ms=1: L.gpudh_temperature(cur, gmax, hbm, mmax, hs, sl, margin, ms, mthr)
    assert T(90, 89, 70, 95, 20, 10) == 1
    assert T(89, 89, 96, 95, 20, 10) == 2
    assert T(80, 89, 70, 95, 10, 10) == 4
    assert T(80, 0, 70, 0, 5, 0) == 0
    assert T(80, 89, 96, 95, 20, 10, hs=0) == 0          # HBM sensor not supported
    assert T(80, 89, 70, 95, 0, 10) == 0 and T(80, 89, 70, 95, -3, 10) == 0     # unreliable margin readings are skipped
    assert T(80, 89, 70, 95, 5, 10, ms=0) == 0 and T(80, 89, 70, 95, 5, 10, sl=0) == 0


# ---- the stateful kmsg matchers (csrc/kmsg_stateful.cpp) over oracle-built primitive hits: no GPU needed ----
def _oracle_prim_hits(lines, unit0=0):
    """XidHit records for the six line primitives, spans from the oracle's regexes (what the scan kernel reports)"""
    import gpud_b200 as g
    from oracle import pyoracle as O
    buf = b"\n".join(lines)
    hits, off = [], 0
    for u, l in enumerate(lines):
        for k in range(19, 25):
            m = O.EXT_RE[k].search(l)
            if not m:
                continue
            h = g.XidHit()
            h.unit_index, h.unit_offset, h.kind, h.link = u + unit0, off, k, off + m.start()
            slots = ("dev", "unit_name", "pid", "pname", "inj")
            for slot, grp in zip(slots, O.PRIM_GROUPS.get(k, (None,) * 5)):
                if grp is not None:
                    setattr(h, slot + "_off", off + m.start(grp))
                    setattr(h, slot + "_len", m.end(grp) - m.start(grp))
            hits.append(h)
        off += len(l) + 1
    return hits, buf


def test_kmsg_stateful_golden_sequences():
    import gpud_b200 as g
    import synth
    from oracle import pyoracle as O
    for seq in synth.stateful_sequences():
        lines = [l.encode() for l in seq]
        hits, buf = _oracle_prim_hits(lines)
        st = g.KmsgStateful()
        assert st.feed(hits, buf, len(lines)) == O.stateful_events(lines), seq
        st.close()


@pytest.mark.parametrize("seed,chunk", [(1, 100000), (2, 7), (3, 1), (4, 64), (5, 13)])
def test_kmsg_stateful_long_stream_in_chunks(seed, chunk):
    """state (panic line counter, OOM instance) carries across feeds: the same events whatever the chunking"""
    import gpud_b200 as g
    import synth
    from oracle import pyoracle as O
    lines = [l.encode() for l in synth.stateful_stream(4000, seed=seed)]
    want = O.stateful_events(lines)
    assert sum(1 for e in want if e[1] == "os") >= 10 and sum(1 for e in want if e[1] == "memory") >= 10
    st = g.KmsgStateful()
    got = []
    for a in range(0, len(lines), chunk):
        part = lines[a:a + chunk]
        hits, buf = _oracle_prim_hits(part)
        got += [(u + a, c, e, m) for u, c, e, m in st.feed(hits, buf, len(part))]
    st.close()
    assert got == want


def test_clock_event_reasons(golden):              # hw-slowdown/clock_events_test.go:21 ; clock_events.go:168-264
    G = golden("clock_events.json")
    tab = G["table"]["entries"]
    assert [(f, h, d) for f, h, d in O.CLOCK_EVENT_REASONS] == [(e["flag"], e["hw_slowdown"], e["description"]) for e in tab]
    L = g.lib()
    hw, other, fl = C.create_string_buffer(4096), C.create_string_buffer(4096), (C.c_int32 * 3)()
    for r in G["reasons"]["rows"]:
        assert O.clock_event_reasons(r["reasons"]) == (r["want_hw"], r["want_other"]), r["name"]
        rc = L.gpud_clock_event_reasons(C.c_uint64(r["reasons"]), hw, 4096, other, 4096, fl)
        assert rc == 100 * len(r["want_hw"]) + len(r["want_other"]), r["name"]
        assert [x for x in hw.value.decode().split("\n") if x] == r["want_hw"] and [x for x in other.value.decode().split("\n") if x] == r["want_other"]
        assert list(fl) == [int(bool(r["reasons"] & 0x8)), int(bool(r["reasons"] & 0x40)), int(bool(r["reasons"] & 0x80))]


ACT_BY_GO = {"IgnoreNoActionRequired": 1, "RebootSystem": 2, "HardwareInspection": 3, "CheckUserAppAndGPU": 4}   # apiv1.RepairActionType* identifiers


def test_sxid_evolve_and_reason(golden):           # sxid/health_state_test.go:38-134 ; sxid/health_state.go:38-111
    L = g.lib()
    out = C.create_string_buffer(256)
    for r in golden("sxid_health.json")["scenarios"]["rows"]:
        ev = [{"k": "xid", "xid": e["code"], "type": e["type"], "actions": e["actions"]} if e["k"] == "err" else {"k": "reboot"} for e in r["events"]]
        arr = (HEvent * max(1, len(ev)))()
        for i, e in enumerate(ev):
            if e["k"] == "xid":
                arr[i].kind, arr[i].event_type, arr[i].xid, arr[i].n_actions = 0, EV[e["type"]], e["xid"], len(e["actions"])
                for k, a in enumerate(e["actions"]):
                    arr[i].actions[k] = ACT_BY_GO[a]
            else:
                arr[i].kind = 1
        h, a, x = C.c_int32(), C.c_int32(), C.c_uint64()
        L.gpudh_evolve_sxid(arr, len(ev), C.byref(h), C.byref(a), C.byref(x))
        oev = [{"name": "error_sxid", "type": e["type"], "xid": e["xid"], "actions": [ACT_BY_GO[q] for q in e["actions"]]} if e["k"] == "xid" else {"name": "reboot"} for e in ev]
        o = O.evolve_healthy_state(oev, 2, "error_sxid")
        assert HEALTH[h.value] == o["health"] and a.value == (o["actions"][0] if o["actions"] else 0), r["name"]
        if "health" in r:
            assert HEALTH[h.value] == r["health"], r["name"]
        if "action" in r:
            assert a.value == (ACT_BY_GO[r["action"]] if r["action"] else 0), r["name"]
        if "reason" in r:
            sx = o["xid"]
            assert O.sxid_reason(sx, "PCI:0000:9b:00") == r["reason"], r["name"]
            n = L.gpud_sxid_reason(C.c_int64(-1 if sx is None else sx), b"PCI:0000:9b:00", out, 256)
            assert n > 0 and out.value.decode() == r["reason"], r["name"]


def test_xid_evolve_extracted_scenarios(golden):   # xid/health_state_test.go:78-246, extracted by script
    for r in golden("xid_health_extracted.json")["scenarios"]["rows"]:
        if not r["events"] and "health" not in r:
            continue
        ev = [{"k": "xid", "xid": e["code"], "type": e["type"], "actions": [a for a in e["actions"]]} if e["k"] == "err" else {"k": "reboot"} for e in r["events"]]
        oev = [{"name": "error_xid", "type": e["type"], "xid": e["xid"], "actions": [ACT_BY_GO[q] for q in e["actions"]]} if e["k"] == "xid" else {"name": "reboot"} for e in ev]
        o = O.evolve_healthy_state(oev, 2)
        if "health" in r:
            assert o["health"] == r["health"], r["name"]
        if r.get("action"):
            assert o["actions"] and o["actions"][0] == ACT_BY_GO[r["action"]], r["name"]


def test_temperature_checks_of_the_reference(golden):   # temperature/component_test.go: TestCheck_* with a Temperature literal
    L = g.lib()
    rows = golden("temperature_checks.json")["checks"]["rows"]
    assert len(rows) >= 10
    for r in rows:
        t = r["temperature"]
        thr = r["margin_threshold"] or 0                 # default 0 = margin rule disabled (temperature/threshold.go:13)
        health, cls = O.temperature_check(t, thr)
        assert health == r["health"], r["name"]
        want_cls = {"margin threshold exceeded": "margin", "GPU temperature anomalies detected": "gpu", "HBM temperature anomalies detected": "hbm",
                    "temperature is": "hbm", "no temperature issue found": ""}[r["reason_contains"][0]]
        assert cls == want_cls, r["name"]
        bits = L.gpudh_temperature(t.get("CurrentCelsiusGPUCore", 0), t.get("ThresholdCelsiusGPUMax", 0), t.get("CurrentCelsiusHBM", 0),
                                   t.get("ThresholdCelsiusMemMax", 0), int(t.get("HBMTemperatureSupported", False)), t.get("ThresholdCelsiusSlowdown", 0),
                                   t.get("ThresholdCelsiusSlowdownMargin", 0), int(t.get("MarginTemperatureSupported", False)), thr)
        got_cls = "margin" if bits & 4 else ("gpu" if bits & 1 else ("hbm" if bits & 2 else ""))
        assert got_cls == cls and (bits != 0) == (health == "Degraded"), r["name"]
