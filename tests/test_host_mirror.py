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
    # the literal calls of Test_parseLine / _WithDifferentBootTimes / _EdgeCases (watcher_test.go:35, :192, :215)
    calls = golden("pkg_kmsg.json")["parse_line_calls"]["rows"]
    assert len(calls) == 6
    for r in calls:
        prio, seq, usec = C.c_int32(), C.c_int64(), C.c_int64()
        msg = C.create_string_buffer(8192)
        assert L.gpudh_parse_kmsg_line(r["input"].encode(), C.byref(prio), C.byref(seq), C.byref(usec), msg, 8192) == 0, r["input"]
        got = {"priority": prio.value, "sequence": seq.value, "usec": usec.value, "message": msg.value.decode()}
        want = dict(zip(("priority", "sequence", "usec", "message"), O.parse_kmsg_line(0, r["input"])))
        assert got == want
        for k in ("priority", "sequence", "usec", "message"):
            if k in r:
                assert got[k] == r[k], (r["input"], k)


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
    assert add(1000 + 901, 1000, "a") == 3 and add(1000 + 901 + 901, 1000, "a") == 1   # live through the last second of its TTL (refreshed by every add), gone after
    L.gpudh_deduper_free(d)
    # the reference's own vectors (pkg/kmsg/deduper_test.go): base = time.Date(2024, 1, 1, 12, 30, 0, 0, time.UTC)
    import datetime
    base = int(datetime.datetime(2024, 1, 1, 12, 30, 0, tzinfo=datetime.timezone.utc).timestamp())
    key = lambda t, m: (L.gpudh_dedup_key(C.c_int64(t), m.encode(), out, 512), out.value.decode())[1]
    want = "%d-%s" % (base - base % 60, "test message")                       # TestCacheKey :137-160: every second of the minute gives the same key
    assert [key(base + s, "test message") for s in (0, 15, 30, 59)] == [want] * 4 == [O.dedup_key(base + s, "test message") for s in (0, 15, 30, 59)]
    k = [key(base + 45, "test message"), key(base + 75, "test message"), key(base + 120, "test message")]     # :162-185 12:30:45, 12:31:15, 12:32:00
    assert len(set(k)) == 3
    assert key(base, "message 1") != key(base + 30, "message 2")              # :187-200
    midnight = int(datetime.datetime(2024, 1, 1, tzinfo=datetime.timezone.utc).timestamp())
    assert key(midnight, "test message") == "%d-test message" % midnight and key(midnight - 1, "test message") != key(midnight, "test message")   # :202-216
    assert key(base + 59, "test message") != key(base + 60, "test message")  # :218-229 last second of a minute vs first of the next
    d = C.c_void_p(L.gpudh_deduper_new(C.c_int64(300)))                      # TestDeduper :63-84 and TestCacheKey :236-268: newDeduper(5 min, ...)
    assert [add(base, base, "test content"), add(base, base + 30, "test content"), add(base, base + 61, "test content")] == [1, 2, 1]
    assert [add(base, base + s, "duplicate message") for s in (0, 30, 59, 60)] == [1, 2, 3, 1]
    assert [add(base, base, "test content 1"), add(base, base, "test content 2"), add(base, base, "test content 1")] == [1, 1, 2]   # :50-61
    L.gpudh_deduper_free(d)
    L.gpudh_deduper_new2.restype = C.c_void_p
    d = C.c_void_p(L.gpudh_deduper_new2(C.c_int64(300), 300))                # :112-133 WithCacheKeyTruncateSeconds(300): +0, +4 min, +5 min
    assert [add(base, base, "test content"), add(base, base + 240, "test content"), add(base, base + 300, "test content")] == [1, 2, 1]
    L.gpudh_deduper_free(d)
    d = C.c_void_p(L.gpudh_deduper_new(C.c_int64(10)))                       # :86-110 expiry: live through its last second, gone after
    assert [add(100, 60, "x"), add(110, 61, "x"), add(121, 62, "x")] == [1, 2, 1]
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
    rows = golden("xid_health_extracted.json")["scenarios"]["rows"]
    assert len(rows) == 10 and all("health" in r for r in rows)
    for r in rows:
        oev = [{"name": "error_xid", "type": e["type"], "xid": e["code"], "actions": None if e["actions"] is None else [ACT_BY_GO[q] for q in e["actions"]]}
               if e["k"] == "err" else {"name": "reboot"} for e in r["events"]]
        o = O.evolve_healthy_state(oev, 2)
        assert o["health"] == r["health"], r["name"]
        if r.get("action"):
            assert o["actions"] and o["actions"][0] == ACT_BY_GO[r["action"]], r["name"]
        if "action" in r and r["action"] is None:
            assert o["actions"] is None
        # the same scenario as STORED events: the payload createXidEvent / createXidEventWithNilSuggestedActions marshals
        # (health_state_test.go:20-39, 59-76), resolved like evolveHealthyState does, on the oracle and on the C++ side
        stored = []
        for e in r["events"]:
            if e["k"] != "err":
                stored.append({"name": "reboot"})
                continue
            data = '{"time":null,"data_source":"test","device_uuid":"PCI:0000:9b:00","xid":%d' % e["code"]
            if e["actions"] is not None:
                data += ',"suggested_actions_by_gpud":{"description":"","repair_actions":["%s"]}' % O.ACTION_WIRE[ACT_BY_GO[e["actions"][0]]]
            stored.append({"name": "error_xid", "type": e["type"], "device_uuid": "PCI:0000:9b:00", "data": data + "}"})
        want = O.evolve_healthy_state_stored(stored, {}, 2)
        assert want["health"] == r["health"], r["name"]
        if r.get("action"):
            assert want["actions"][0] == ACT_BY_GO[r["action"]], r["name"]
        if "reason" in r:
            assert want["reason"] == r["reason"], r["name"]
        assert _evolve_stored(stored, {}, 2) == (want["health"], (want["actions"] or [0])[0], want["reason"]), r["name"]
    # "invalid xid" (:236-246): an undecodable payload is skipped -> Healthy
    bad = [{"name": "error_xid", "type": "Critical", "device_uuid": "", "data": "invalid json"}]
    assert O.evolve_healthy_state_stored(bad, {}, 2)["health"] == "Healthy" and _evolve_stored(bad, {}, 2)[0] == "Healthy"


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
        # the public entry over the reading a poller returns
        pt = g.Temperature()
        pt.current_gpu_core_c, pt.threshold_gpu_max_c, pt.current_hbm_c = t.get("CurrentCelsiusGPUCore", 0), t.get("ThresholdCelsiusGPUMax", 0), t.get("CurrentCelsiusHBM", 0)
        pt.threshold_mem_max_c, pt.hbm_supported, pt.threshold_slowdown_c = t.get("ThresholdCelsiusMemMax", 0), int(t.get("HBMTemperatureSupported", False)), t.get("ThresholdCelsiusSlowdown", 0)
        pt.slowdown_margin_c, pt.margin_supported = t.get("ThresholdCelsiusSlowdownMargin", 0), int(t.get("MarginTemperatureSupported", False))
        assert g.temperature_check(pt, thr) == bits, r["name"]
        # the component's health + reason for a one-GPU box, with the fragments the reference's test asserts
        want_h, want_reason = O.temperature_reason([t], ["GPU-12345678"], thr)
        got_h, got_reason = g.temperature_reason([pt], ["GPU-12345678"], thr)
        assert (["Healthy", "Degraded"][got_h], got_reason) == (want_h, want_reason) and want_h == r["health"], r["name"]
        for frag in r["reason_contains"]:
            assert frag in want_reason, (r["name"], frag, want_reason)


# ---- xid event message / health-state reason (xid/health_state.go:130-281; health_state_test.go:299-942) --------------------
def _go_payload(ev):
    """json.Marshal(xidErrorEventDetail) as createXidEvent / createNVLinkXidEvent build it (health_state_test.go:20-39, 887-905)"""
    import json
    parts = ['"time":null', '"data_source":"test"', '"device_uuid":%s' % json.dumps(ev["device_uuid"]), '"xid":%d' % ev["xid"]]
    if ev["sub_code"]:
        parts.append('"sub_code":%d' % ev["sub_code"])
    if ev["error_status"]:
        parts.append('"error_status":%d' % ev["error_status"])
    if ev["description"]:
        parts.append('"description":%s' % json.dumps(ev["description"]))
    wire = {"RebootSystem": "REBOOT_SYSTEM", "HardwareInspection": "HARDWARE_INSPECTION", "CheckUserAppAndGPU": "CHECK_USER_APP_AND_GPU", "IgnoreNoActionRequired": "IGNORE_NO_ACTION_REQUIRED"}
    parts.append('"suggested_actions_by_gpud":{"description":"","repair_actions":["%s"]}' % wire[ev["action"]])
    return "{" + ",".join(parts) + "}"


def _devspec(devices):
    return ";".join("%s=%s" % kv for kv in devices.items()).encode()


def _evolve_stored(events, devices, thr=2):
    L = g.lib()
    blob = "\x1e".join("\x1f".join([e["name"], e.get("type", ""), e.get("device_uuid", ""), e.get("data", "")]) for e in events)
    h, a = C.c_int32(), C.c_int32()
    out = C.create_string_buffer(2048)
    L.gpudh_evolve_stored(blob.encode("utf-8"), _devspec(devices), thr, C.byref(h), C.byref(a), out, 2048)
    return HEALTH[h.value], a.value, out.value.decode("utf-8")


def test_xid_build_message_vectors_of_the_reference(golden):
    G = golden("xid_messages.json")
    assert len(G["literals"]["rows"]) == 7 and len(G["standard"]["rows"]) == 6
    for r in G["literals"]["rows"]:                                  # xidErrorEventDetail literals + asserts on buildMessage(nil)
        f = r["fields"]
        args = (f["Xid"], f.get("SubCode", 0), f.get("ErrorStatus", 0), f.get("Description", ""), f.get("DeviceUUID", ""))
        want = O.xid_build_message(*args)
        assert g.xid_build_message(*args) == want
        if r["equal"] is not None:
            assert want == r["equal"], r["func"]
        for c in r["contains"]:
            assert c in want, (r["func"], c)
    S = G["standard"]
    for r in S["rows"]:                                              # Test_HealthStateReason_StandardXIDs: devices {uuid: bus id}
        uuid = O.convert_bus_id_to_uuid(r["device_uuid"], {S["uuid"]: S["bus_id"]})
        assert uuid == S["uuid"] and g.lib().gpud_xid_device_matches_bus_id(r["device_uuid"].encode(), S["bus_id"].encode()) == 1
        want = O.xid_build_message(r["xid"], 0, 0, r["description"], r["device_uuid"], uuid)
        assert g.xid_build_message(r["xid"], 0, 0, r["description"], r["device_uuid"], uuid) == want
        for c in r["contains"]:
            assert c in want, (r["name"], c)
        assert ("XID %d." % r["xid"]) not in want and "err status" not in want
    assert g.lib().gpud_xid_device_matches_bus_id(b"PCI:0000:04:00", b"0000:9b:00.0") == 0
    assert g.lib().gpud_xid_device_matches_bus_id(b"0000:04:00", b"0000:04:00.0") == 1          # TrimPrefix: the "PCI:" is optional
    # the uint64 that does not fit an int (intFromUint64 fails): header and device only
    assert g.xid_build_message(2 ** 63, 0, 0, "x", "PCI:1") == O.xid_build_message(2 ** 63, 0, 0, "x", "PCI:1") == "XID 9223372036854775808 detected on GPU PCI:1"
    # "Unused" and a description equal to the mnemonic are not appended
    for xid in (1, 13, 79, 94, 144, 150, 173, 999):
        for desc in ("", "Unused", O.MNEMONIC.get(xid, ""), "something else"):
            assert g.xid_build_message(xid, 3, 0x10, desc, "PCI:0000:01:00", "GPU-u") == O.xid_build_message(xid, 3, 0x10, desc, "PCI:0000:01:00", "GPU-u")


def test_xid_reason_from_kmsg_lines_oracle(golden):
    G = golden("xid_messages.json")["from_lines"]
    assert len(G["rows"]) == 15
    for r in G["rows"]:
        x = O.xid_match(r["line"].encode())
        assert x is not None and x.detail is not None, r["name"]
        d = x.detail
        msg = O.xid_build_message(x.xid, d.sub_code, d.error_status, d.description, x.device, O.convert_bus_id_to_uuid(x.device, r["devices"]))
        for c in r["contains"]:
            assert c in msg, (r["name"], c, msg)
        if "sub_code" in r:
            assert d.sub_code == r["sub_code"] and ("%d.%d" % (x.xid, r["sub_code"])) in msg
        if "xid" in r:
            assert x.xid == r["xid"]
        if "event_type" in r:
            assert O.EVENT_NAMES[d.event_type] == r["event_type"]
        if "hint" in r:
            assert d.investigatory_hint == r["hint"]
        # gpud_xid_hit_message over a hit carrying what the scan fills in (the scan itself is compared on the GPU)
        h = g.XidHit()
        h.kind, h.code, h.device, h.dev_len = 1, x.xid, x.device.encode(), len(x.device)
        h.flags, h.sub_code, h.error_status = (1 if x.info is not None and x.info.unit else 0), d.sub_code, d.error_status
        L = g.lib()
        L.gpud_xid_description.restype = C.c_char_p
        h.detail_variant = next(v for v in (0, 1, 2) if L.gpud_xid_description(x.xid, v).decode() == d.description)
        uuid = O.convert_bus_id_to_uuid(x.device, r["devices"])
        assert g.xid_hit_message(h, uuid) == msg, r["name"]


def test_evolve_over_stored_events_with_reason(golden):
    G = golden("xid_messages.json")["evolve"]
    assert len(G["rows"]) == 5
    for r in G["rows"]:                                              # Test_HealthStateReason_evolveHealthyState_Integration
        events = [dict(e, data=_go_payload(e)) if e["name"] == "error_xid" else e for e in r["events"]]
        want = O.evolve_healthy_state_stored(events, G["devices"])
        assert want["health"] == r["health"], r["name"]
        for c in r["contains"]:
            assert c in want["reason"], (r["name"], c)
        got = _evolve_stored(events, G["devices"])
        assert got == (want["health"], (want["actions"] or [0])[0], want["reason"]), r["name"]

def test_resolve_xid_event_matches_the_oracle_on_random_payloads():
    import json
    import numpy as np
    L = g.lib()
    rng = np.random.default_rng(5)
    devices = {"GPU-aaaa": "0000:04:00.0", "GPU-bbbb": "0000:9b:00.0"}
    wires = list(O.ACTION_WIRE.values())
    codes = [0, 1, 13, 31, 63, 64, 79, 94, 119, 123, 144, 145, 149, 150, 154, 172, 173, 999, 99999]
    n_res = 0
    for i in range(3000):
        kind = rng.random()
        typ = str(rng.choice(["", "Warning", "Critical", "Fatal", "Info"]))
        dev = str(rng.choice(["PCI:0000:04:00", "PCI:0000:9b:00", "PCI:0000:01:00", ""]))
        if kind < 0.15:                                              # legacy rows: a decimal code, or junk
            raw = str(rng.choice(["79", "94", "+63", "99999", "0", "-5", "12x", "", " 79", "149", "1e3", "0079"]))
        else:
            xid = int(rng.choice(codes))
            p = {"time": None, "data_source": "kmsg", "device_uuid": dev, "xid": xid}
            if rng.random() < 0.5:
                p["sub_code"] = int(rng.choice([0, 4, 10, 37, 38, 63]))
            if rng.random() < 0.5:
                p["error_status"] = int(rng.choice([0, 1, 2, 8, 0x80000000]))
            if rng.random() < 0.5:
                p["description"] = str(rng.choice(["", "Unused", "some \"quoted\" text", "NVLINK: RLW Error", "café \U0001F600"]))
            if rng.random() < 0.3:
                p["investigatory_hint"] = "INVESTIGATE_PEER_DEVICE"
            r = rng.random()
            if r < 0.5:
                p["suggested_actions_by_gpud"] = {"repair_actions": [str(a) for a in rng.choice(wires, size=int(rng.integers(0, 3)))]}
            elif r < 0.6:
                p["suggested_actions_by_gpud"] = None
            raw = json.dumps(p, separators=(",", ":"), ensure_ascii=bool(rng.random() < 0.5))
            if kind > 0.95:
                raw = raw[:-3]                                       # truncated JSON
        want = O.resolve_xid_event(typ, raw, dev, devices)
        tout, mout = C.create_string_buffer(64), C.create_string_buffer(2048)
        na, acts = C.c_int32(), (C.c_int32 * 4)()
        ok = L.gpudh_resolve_xid_event(typ.encode(), raw.encode("utf-8"), dev.encode(), _devspec(devices), tout, 64, mout, 2048, C.byref(na), acts)
        assert bool(ok) == (want is not None), (typ, raw)
        if want is None:
            continue
        n_res += 1
        wt, wm, wp = want
        assert (tout.value.decode(), mout.value.decode("utf-8")) == (wt, wm), raw
        assert (None if na.value < 0 else [acts[k] for k in range(na.value)]) == wp["actions"], raw
    assert n_res > 2000
    # and the fold over random stored histories
    for i in range(300):
        events = []
        for _ in range(int(rng.integers(0, 9))):
            if rng.random() < 0.3:
                events.append({"name": "reboot"})
            else:
                xid = int(rng.choice([13, 31, 79, 94, 123, 149, 150, 99999]))
                ev = {"name": "error_xid", "type": str(rng.choice(["Warning", "Critical", "Fatal"])), "xid": xid, "sub_code": int(rng.choice([0, 4, 37])),
                      "error_status": int(rng.choice([0, 2])), "device_uuid": "PCI:0000:04:00", "description": str(rng.choice(["", "d"])),
                      "action": str(rng.choice(["RebootSystem", "HardwareInspection", "CheckUserAppAndGPU"]))}
                ev["data"] = _go_payload(ev) if rng.random() < 0.85 else str(xid)
                events.append(ev)
        thr = int(rng.integers(1, 4))
        want = O.evolve_healthy_state_stored(events, devices, thr)
        assert _evolve_stored(events, devices, thr) == (want["health"], (want["actions"] or [0])[0], want["reason"]), events


def test_temperature_reason_over_a_box():
    import numpy as np
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(0, 9))
        temps, structs, uuids = [], [], []
        for i in range(n):
            t = {"CurrentCelsiusGPUCore": int(rng.integers(30, 100)), "CurrentCelsiusHBM": int(rng.integers(30, 110)), "HBMTemperatureSupported": bool(rng.random() < 0.7),
                 "ThresholdCelsiusSlowdown": int(rng.choice([0, 95, 100])), "ThresholdCelsiusMemMax": int(rng.choice([0, 95, 105])),
                 "ThresholdCelsiusGPUMax": int(rng.choice([0, 85, 88])), "ThresholdCelsiusSlowdownMargin": int(rng.integers(-5, 60)),
                 "MarginTemperatureSupported": bool(rng.random() < 0.7)}
            pt = g.Temperature()
            pt.current_gpu_core_c, pt.current_hbm_c, pt.hbm_supported = t["CurrentCelsiusGPUCore"], t["CurrentCelsiusHBM"], int(t["HBMTemperatureSupported"])
            pt.threshold_slowdown_c, pt.threshold_mem_max_c, pt.threshold_gpu_max_c = t["ThresholdCelsiusSlowdown"], t["ThresholdCelsiusMemMax"], t["ThresholdCelsiusGPUMax"]
            pt.slowdown_margin_c, pt.margin_supported = t["ThresholdCelsiusSlowdownMargin"], int(t["MarginTemperatureSupported"])
            temps.append(t); structs.append(pt); uuids.append("GPU-%04d" % i)
        thr = int(rng.choice([0, 5, 10, 30]))
        want_h, want_reason = O.temperature_reason(temps, uuids, thr)
        got_h, got_reason = g.temperature_reason(structs, uuids, thr)
        assert (["Healthy", "Degraded"][got_h], got_reason) == (want_h, want_reason)


def test_product_capabilities_of_the_reference(golden):   # pkg/nvidia/product/capabilities_test.go:8, :168, :230
    L = g.lib()
    G = golden("product_caps.json")
    assert (len(G["mem_caps"]["rows"]), len(G["fm_supported"]["rows"]), len(G["fabric_state_supported"]["rows"])) == (18, 8, 19)
    for r in G["mem_caps"]["rows"]:
        assert L.gpud_product_mem_caps(r["product"].encode()) == O.product_mem_caps(r["product"]) == r["caps"], r["name"]
    for r in G["fm_supported"]["rows"]:
        assert bool(L.gpud_product_fm_supported(r["product"].encode())) == O.product_fm_supported(r["product"]) == r["expected"], r["name"]
    for r in G["fabric_state_supported"]["rows"]:
        assert bool(L.gpud_product_fabric_state_supported(r["product"].encode())) == O.product_fabric_state_supported(r["product"]) == r["expected"], r["name"]
    for name in ("NVIDIA B200", "NVIDIA GB200 NVL72", "NVIDIA GH200 480GB", "NVIDIA H200 PCIe", "NVIDIA A100-SXM4-80GB", "NVIDIA A10G", "Tesla T4", ""):
        assert L.gpud_product_mem_caps(name.encode()) == O.product_mem_caps(name)
        assert bool(L.gpud_product_fm_supported(name.encode())) == O.product_fm_supported(name)
        assert bool(L.gpud_product_fabric_state_supported(name.encode())) == O.product_fabric_state_supported(name)
    assert L.gpud_product_mem_caps(b"NVIDIA B200") == 7 and L.gpud_product_fm_supported(b"NVIDIA B200") == 1      # this framework's target


# ---- sxid stored events: resolveSXIDEvent + evolveHealthyState (sxid/health_state.go:38-142) ---------------------------------
def _evolve_stored_sxid(events):
    L = g.lib()
    blob = "\x1e".join("\x1f".join([e["name"], e.get("type", ""), e.get("device_uuid", ""), e.get("data", "")]) for e in events)
    h, a = C.c_int32(), C.c_int32()
    out = C.create_string_buffer(1024)
    L.gpudh_evolve_stored_sxid(blob.encode("utf-8"), C.byref(h), C.byref(a), out, 1024)
    return HEALTH[h.value], a.value, out.value.decode("utf-8")


def test_sxid_stored_events_resolve_and_evolve(golden):
    import numpy as np
    L = g.lib()
    # catalog accessor == the oracle's table
    for code, d in O.SXID_DETAILS.items():
        ev, na = C.c_int32(), C.c_int32()
        acts = (C.c_int32 * 4)()
        assert L.gpud_sxid_get_detail(code, C.byref(ev), C.byref(na), acts) == 1
        assert ev.value == d["event_type"] and (None if na.value < 0 else [acts[i] for i in range(na.value)]) == (list(d["actions"]) if d["actions"] else None), code
    assert L.gpud_sxid_get_detail(99999, None, None, None) == 0
    # the scenarios of sxid/health_state_test.go with the payload createSXidEvent marshals (:15-35): already JSON, so the event's own
    # type and actions count and the reason has no catalog name for the made-up codes
    for r in golden("sxid_health.json")["scenarios"]["rows"]:
        stored = []
        for e in r["events"]:
            if e["k"] != "err":
                stored.append({"name": "reboot"})
                continue
            data = '{"time":null,"data_source":"test","device_uuid":"PCI:0000:9b:00","sxid":%d,"suggested_actions_by_gpud":{"description":"","repair_actions":["%s"]}}' % (
                e["code"], O.ACTION_WIRE[ACT_BY_GO[e["actions"][0]]])
            stored.append({"name": "error_sxid", "type": e["type"], "device_uuid": "", "data": data})
        want = O.evolve_sxid_stored(stored)
        if "health" in r:
            assert want["health"] == r["health"], r["name"]
        if r.get("action"):
            assert want["actions"][0] == ACT_BY_GO[r["action"]], r["name"]
        if "reason" in r:
            assert want["reason"] == r["reason"], r["name"]
        assert _evolve_stored_sxid(stored) == (want["health"], (want["actions"] or [0])[0], want["reason"]), r["name"]
    # what the component itself persists: the decimal code (sxid/component.go:448-455), resolved from the catalog on read
    rng = np.random.default_rng(9)
    codes = list(O.SXID_DETAILS)[:40] + [99999, 0]
    for i in range(400):
        events = []
        for _ in range(int(rng.integers(0, 8))):
            if rng.random() < 0.3:
                events.append({"name": "reboot"})
            else:
                code = int(rng.choice(codes))
                raw = str(rng.choice([str(code), str(code), "+%d" % code, "-%d" % code, "x%d" % code, '{"sxid":%d,"device_uuid":"PCI:0000:0%d:00"}' % (code, i % 8), "{bad"]))
                events.append({"name": "error_sxid", "type": str(rng.choice(["", "Warning", "Fatal"])), "device_uuid": "PCI:0000:0%d:00" % (i % 8), "data": raw})
        want = O.evolve_sxid_stored(events)
        assert _evolve_stored_sxid(events) == (want["health"], (want["actions"] or [0])[0], want["reason"]), events


def test_nvml_error_classes_of_the_reference(golden):   # pkg/nvidia/errors/error_test.go:11, :451, :535 (constant rows + mocked error strings)
    L = g.lib()
    G = golden("nvml_error_classes.json")
    assert {k: len(v["rows"]) for k, v in G.items()} == {"not_supported": 12, "gpu_lost": 7, "reset_required": 9}
    # what nvmlErrorString / go-nvml's fallback say for the constants the tables use
    canon = {0: ("Success", "SUCCESS"), 3: ("Not Supported", "ERROR_NOT_SUPPORTED"), 15: ("GPU is lost", "ERROR_GPU_IS_LOST"), 16: ("GPU requires reset", "ERROR_RESET_REQUIRED"),
             25: ("Argument version mismatch", "ERROR_ARGUMENT_VERSION_MISMATCH"), 999: ("Unknown Error", "ERROR_UNKNOWN"), 27: ("Not Ready", "ERROR_NOT_READY"),
             6: ("Not Found", "ERROR_NOT_FOUND")}
    bit = {"not_supported": 1, "gpu_lost": 2, "reset_required": 4}
    for key, tab in G.items():
        for r in tab["rows"]:
            texts = [r["error_string"]] if r["error_string"] is not None else list(canon[r["ret"]])
            for text in texts:
                got = L.gpudh_nvml_error_class(r["ret"], text.encode())
                assert bool(got & bit[key]) == r["expected"], (key, r["name"], text)


def test_catalog_detail_lookups_on_the_references_tables(golden):
    """TestGetDetailWithSubCode (xid/kmsg_extended_test.go:356-464) through gpud_xid_detail with a status no rule carries (so the lookup is the
    sub-code fallback chain), Test_detailFromNVLinkInfo_StatusSpecific's codes through the status table, GetDetail through gpud_xid_get_detail"""
    L = g.lib()
    G = golden("xid_kmsg.json")
    assert all(r["ErrorStatus"] != 0xFFFFFFFE for r in O.NVLINK_RULES)
    for r in G["detail_with_subcode"]["rows"]:
        d = g.xid_detail(r["xid"], r["subCode"], 0xFFFFFFFE)
        o = O.get_detail_with_sub_code(r["xid"], r["subCode"])
        assert (d is not None) == r["expectedFound"] == (o is not None), r["name"]
        if d is None:
            continue
        if r["expectedEventTypeFatal"]:
            assert d["event_type"] == O.EV_FATAL, r["name"]
        assert (d["event_type"], d["actions"], d["description"], d["sub_code"]) == (o.event_type, o.actions, o.description, o.sub_code), r["name"]
    for code in list(range(0, 180)) + [99999, -3]:
        ev, na = C.c_int32(), C.c_int32()
        acts = (C.c_int32 * 4)()
        found = L.gpud_xid_get_detail(code, C.byref(ev), C.byref(na), acts)
        o = O.get_detail(code)
        assert bool(found) == (o is not None), code
        if o is not None:
            assert (ev.value, None if na.value < 0 else [acts[i] for i in range(na.value)]) == (o.event_type, o.actions), code
    assert sum(1 for c in range(0, 200) if L.gpud_xid_get_detail(c, None, None, None)) == 172          # {1..173} \\ {133}


def test_merge_and_trim_scenarios_of_the_reference():
    """TestMergeEvents / TestTrimEventsAfterSetHealthy (xid/component_test.go:81-207; the sxid twins are the same functions)"""
    L = g.lib()
    now = 1_770_000_000
    H, M = 3600, 60

    def merge(a, b):
        aa, bb = (C.c_int64 * max(1, len(a)))(*a), (C.c_int64 * max(1, len(b)))(*b)
        out = (C.c_int64 * max(1, len(a) + len(b)))()
        n = L.gpudh_merge_times(aa, len(a), bb, len(b), out)
        return [out[i] for i in range(n)]
    for a, b, want_len in (([], [], 0), ([], [now], 1), ([now], [], 1), ([now - H, now], [now - 2 * H, now - 30 * M], 4)):     # the table
        got = merge(a, b)
        assert len(got) == want_len and got == sorted(got, reverse=True)
    assert merge([now + 2 * H, now - H], [now, now - 2 * H]) == [now + 2 * H, now, now - H, now - 2 * H]                           # "verify sorting"
    assert L.gpudh_trim_count(b"error_xid\nSetHealthy\nerror_xid") == 1          # dropsEventsOlderThanSetHealthy
    assert L.gpudh_trim_count(b"error_xid\nerror_xid") == 2                       # noSetHealthyReturnsOriginal
    assert L.gpudh_trim_count(b"SetHealthy") == 0                                 # onlySetHealthyReturnsEmpty
    assert O.trim_events_after_set_healthy([{"name": "error_xid"}, {"name": "SetHealthy"}, {"name": "error_xid"}]) == [{"name": "error_xid"}]
    assert O.trim_events_after_set_healthy([{"name": "SetHealthy"}]) == []


def test_poll_row_holds_the_last_good_value_when_a_getter_fails():
    """oracle/SPEC.md: a failed getter repeats the column's last good value (0 before the first good read); no sentinel ever
    reaches the ring.  NVML return codes: 0 success, 3 not supported, 999 unknown (transient)."""
    held = [0] * 8
    row, mask = g.capi.poll_row_hold([41, 250000, 1965, 1965, 3996, 97, 30, 1024], [0] * 8, held)
    assert row == [41, 250000, 1965, 1965, 3996, 97, 30, 1024] and mask == 0
    # a transient failure of the power getter and an unsupported memory-utilisation getter: both columns repeat
    row, mask = g.capi.poll_row_hold([42, 0xDEAD, 1950, 1950, 3996, 98, 0xBEEF, 1030], [0, 999, 0, 0, 0, 0, 3, 0], held)
    assert row == [42, 250000, 1950, 1950, 3996, 98, 30, 1030] and mask == (1 << 1) | (1 << 6)
    assert 0xFFFFFFFF not in row
    # the column recovers with the next good read
    row, mask = g.capi.poll_row_hold([43, 260000, 1950, 1950, 3996, 99, 31, 1031], [0] * 8, held)
    assert row == [43, 260000, 1950, 1950, 3996, 99, 31, 1031] and mask == 0
    # never read successfully: the column stays 0, not 0xffffffff
    held2 = [0] * 8
    row, mask = g.capi.poll_row_hold([7] * 8, [3] * 8, held2)
    assert row == [0] * 8 and mask == 0xFF


def test_remapped_rows_check_of_the_reference():
    """remapped-rows/component.go:197-300 with the devices of TestCheckOnceRemappingIssueDetection (component_test.go:405-517) and
    TestCheckOnceWithMultipleGPUs (:1288-1394): reason fragments, health and the suggested action (RMA takes precedence)."""
    R = g.capi.RemappedRows
    rows = [R(0, 0, 0, 0, 1, 0), R(0, 0, 1, 0, 1, 0), R(0, 2, 0, 1, 1, 0)]
    health, action, reason = g.capi.remapped_rows_check(rows, ["0000:01:00.0", "0000:02:00.0", "0000:03:00.0"])
    assert health == 2 and action == 3                      # Unhealthy, HARDWARE_INSPECTION
    assert reason == "0000:02:00.0 needs reset (detected pending row remapping), 0000:03:00.0 qualifies for RMA (row remapping failed, remapped due to 2 uncorrectable error(s))"
    # the order of the devices does not let a later pending GPU take the action back from an earlier failed one
    health, action, reason = g.capi.remapped_rows_check(rows[::-1], ["0000:03:00.0", "0000:02:00.0", "0000:01:00.0"])
    assert health == 2 and action == 3 and reason.startswith("0000:03:00.0 qualifies for RMA")
    # only pending: reboot
    health, action, reason = g.capi.remapped_rows_check(rows[:2], ["0000:01:00.0", "0000:02:00.0"])
    assert (health, action, reason) == (2, 2, "0000:02:00.0 needs reset (detected pending row remapping)")
    # a GPU that is both failed and pending reports both issues, RMA first (:288-293)
    health, action, reason = g.capi.remapped_rows_check([R(1, 9, 1, 1, 1, 0)], ["0000:0f:00.0"])
    assert action == 3 and reason == ("0000:0f:00.0 qualifies for RMA (row remapping failed, remapped due to 9 uncorrectable error(s)), "
                                      "0000:0f:00.0 needs reset (detected pending row remapping)")
    # nothing wrong
    assert g.capi.remapped_rows_check(rows[:1] * 8, ["x"] * 8) == (0, 0, "8 devices support remapped rows and found no issue")
    assert g.capi.remapped_rows_check([], []) == (0, 0, "0 devices support remapped rows and found no issue")


def test_nvml_bus_id_like_go_nvlib():
    """go-nvlib device.GetPCIBusID: lower case, the first four zeros of NVML's eight-digit domain dropped"""
    assert g.capi.nvml_bus_id("00000000:3B:00.0") == "0000:3b:00.0"
    assert g.capi.nvml_bus_id("00000001:0F:00.0") == "0001:0f:00.0"
    assert g.capi.nvml_bus_id("0000:9b:00.0") == ":9b:00.0"          # what TrimPrefix does to an already short id
    assert g.capi.nvml_bus_id("0000") == "0000"


def test_watcher_dedup_in_front_of_the_stateful_matchers():
    """pkg/kmsg/watcher.go:281-286: a message is handed to the matchers only the first time its (minute, message) key shows up, so a
    repeated line neither matches nor counts towards the panic matcher's ten lines (os/kmsg_matcher.go:60-125).  Model: the oracle's
    matchers run over the stream with the duplicates removed by the reference's deduper restatement (O.Deduper)."""
    import synth
    base = [l for l in synth.stateful_stream(1500, seed=11)]
    # repeat lines inside and outside the panic windows: every third line once more, right away and again five lines later
    lines = []
    for i, l in enumerate(base):
        lines.append(l)
        if i % 3 == 0:
            lines.append(l)
        if i % 5 == 0 and i >= 5:
            lines.append(base[i - 5])
    enc = [l.encode() for l in lines]
    buf = b"\n".join(enc)
    now = 1_700_000_000
    # the reference model: drop duplicates with the oracle's deduper, run the closures over what is left, report original line numbers
    seen = set()                                       # deduper.addCache (deduper.go:95-109): the first add of a key returns 1, later ones > 1
    kept = []
    for i, l in enumerate(lines):
        k = O.dedup_key(now, l)
        if k not in seen:
            seen.add(k)
            kept.append(i)
    want = [(kept[u], c, e, m) for u, c, e, m in O.stateful_events([enc[i] for i in kept])]
    assert len(kept) < len(lines) * 0.8 and sum(1 for w in want if w[1] == "os") >= 5
    dd = g.capi.KmsgDeduper()
    dropped = dd.units(buf, len(enc), mode=g.SCAN_LINES, lines_unix=now, now_unix=now)
    assert [i for i in range(len(enc)) if not dropped[i]] == kept
    hits, hbuf = _oracle_prim_hits(enc)
    assert hbuf == buf
    st = g.KmsgStateful()
    got = st.feed_units(hits, buf, len(enc), dropped)
    assert got == want
    # without the dedup the repeated lines shift the ten-line window: the results differ (what the gap used to be)
    st2 = g.KmsgStateful()
    assert st2.feed(hits, buf, len(enc)) != want
    st.close(); st2.close(); dd.close()
    # RAW_KMSG: the key's minute comes from the record's own timestamp - the same message a minute later is not a duplicate
    recs = [b"6,1,1000000,-;hello", b"6,2,2000000,-;hello", b"6,3,62000000,-;hello", b"bad record", b"6,4,63000000,-;hello\n SUBSYSTEM=x", b"6,5,64000000,-;hello\n SUBSYSTEM=x"]
    dd = g.capi.KmsgDeduper()
    dr = dd.units(b"\n".join(recs), len(recs), mode=g.SCAN_RAW_KMSG, boot_unix=now - now % 60, now_unix=now)
    assert list(dr) == [0, 1, 0, 0, 0, 1]
    dd.close()


def test_gpm_check_of_the_reference():
    """gpm/component.go:196-290: one GPU without GPM -> Healthy "GPM not supported"; else the all-checked reason"""
    M = g.capi.GpmMetrics
    on, off = M(), M()
    on.supported = 1
    assert g.capi.gpm_check([on, on, on]) == (0, "all 3 GPU(s) were checked, no GPM issue found")
    assert g.capi.gpm_check([on, off, on]) == (0, "GPM not supported")
    assert g.capi.gpm_check([]) == (0, "all 0 GPU(s) were checked, no GPM issue found")
    assert g.capi.GPM_METRIC_IDS == (3, 4, 5, 6, 7, 9, 11, 12, 13) and len(g.capi.GPM_METRICS) == 9
