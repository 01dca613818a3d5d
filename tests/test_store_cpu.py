"""The SQLite write path (gpud_store_*, csrc/store_sqlite.cpp) against the reference's own SQL (tests/golden/store_sql.json,
extracted from pkg/eventstore/database.go and pkg/metrics/store/sqlite.go): same DDL, same rows, and the reference's
reader query returns them.  CPU only."""
import json
import re
import sqlite3

import pytest

import gpud_b200 as g


def _norm(sql):
    return re.sub(r"\s+", " ", sql).strip().rstrip(";")


def _go_table_name(component, version):            # defaultTableName (pkg/eventstore/database.go:136-143)
    c = component.replace(" ", "_").replace("-", "_").replace("__", "_").lower()
    return "components_%s_events_%s" % (c, version)


@pytest.fixture
def store(tmp_path):
    try:
        st = g.Store(str(tmp_path / "gpud.state"))
    except g.GpudError as e:
        pytest.skip("no libsqlite3.so.0: %s" % e)
    yield st, str(tmp_path / "gpud.state")
    st.close()


def test_event_table_is_the_references(store, golden):
    st, path = store
    G = golden("store_sql.json")
    ver = G["constants"]["event_schema_version"]
    for r in G["table_names"]["rows"]:                  # Test_defaultTableName (pkg/eventstore/database_test.go:19)
        assert st.event_table(r["input"]) == r["expected"], r["name"]
    for comp in ("accelerator-nvidia-error-xid", "accelerator-nvidia-error-sxid", "os", "Some Comp--Name", "a__b___c", "x - y"):
        t = st.event_table(comp)
        assert t == _go_table_name(comp, ver)
        db = sqlite3.connect(path)
        got = {r[0]: r[1] for r in db.execute("SELECT name, sql FROM sqlite_master WHERE tbl_name = ?", (t,))}
        assert _norm(got[t]) == _norm(G["event_create_table"]["sql"].format(table=t).replace("IF NOT EXISTS ", ""))
        for k in ("event_index_0", "event_index_1", "event_index_2"):
            want = G[k]["sql"].format(table=t)
            name = re.search(r"(idx_\w+) ON", want).group(1)
            assert _norm(got[name]) == _norm(want.replace("IF NOT EXISTS ", "")), k
        db.close()


def test_insert_and_read_back_with_the_references_query(store, golden):
    st, path = store
    G = golden("store_sql.json")
    t = st.event_table("accelerator-nvidia-error-xid")
    extra = json.dumps({"data": "{\"xid\":79}", "device_uuid": "PCI:0000:05:00"}, separators=(",", ":"), sort_keys=True)
    st.insert_event(t, 1740000000, "error_xid", "Fatal", "", extra)
    st.insert_event(t, 1740000100, "reboot", "Warning", "system reboot detected", "")
    st.insert_event(t, 1739999000, "error_xid", "Warning", "", extra)
    db = sqlite3.connect(path)
    rows = list(db.execute(G["event_get"]["sql"].format(table=t), (1739999500,)))
    assert rows == [(1740000100, "reboot", "Warning", "system reboot detected", None),     # NULLIF('', '') -> NULL, newest first
                    (1740000000, "error_xid", "Fatal", None, extra)]
    # the reference's own INSERT statement yields the same row shape
    db.execute(G["event_insert"]["sql"].format(table=t), (1740000200, "error_xid", "Info", "", extra))
    db.commit()
    a = list(db.execute("SELECT typeof(message), typeof(extra_info) FROM %s WHERE timestamp IN (1740000000, 1740000200) ORDER BY timestamp" % t))
    assert a[0] == a[1] == ("null", "text")
    db.close()


def test_metrics_store(store, golden):
    st, path = store
    G = golden("store_sql.json")
    ver = G["constants"]["metrics_schema_version"]
    st.metrics_table()
    t = "gpud_metrics_%s" % ver
    db = sqlite3.connect(path)
    (sql,) = db.execute("SELECT sql FROM sqlite_master WHERE name = ?", (t,)).fetchone()
    assert _norm(sql) == _norm(G["metrics_create_table"]["sql"].format(table=t).replace("IF NOT EXISTS ", ""))
    labels = json.dumps({"gpu": "GPU-0", "window": "17"}, separators=(",", ":"), sort_keys=True)
    st.record_metrics([(1740000000123, "accelerator-nvidia-temperature", "window_p99_celsius", labels, 71.5),
                       (1740000000123, "accelerator-nvidia-temperature", "window_mean_celsius", labels, 64.25),
                       (1740000000123, "accelerator-nvidia-power", "window_max_milliwatts", "", 412000.0)])
    st.record_metrics([(1740000000123, "accelerator-nvidia-temperature", "window_p99_celsius", labels, 72.0)])    # same key: OR REPLACE
    rows = list(db.execute("SELECT unix_milliseconds, component_name, metric_name, metric_labels, metric_value FROM %s ORDER BY unix_milliseconds ASC, metric_name" % t))
    assert rows == [(1740000000123, "accelerator-nvidia-power", "window_max_milliwatts", "", 412000.0),
                    (1740000000123, "accelerator-nvidia-temperature", "window_mean_celsius", labels, 64.25),
                    (1740000000123, "accelerator-nvidia-temperature", "window_p99_celsius", labels, 72.0)]
    # the reference's own multi-row insert lands in the same table
    db.execute(G["metrics_insert_prefix"]["sql"].format(table=t) + "(?, ?, ?, ?, ?)", (1740000000999, "c", "m", "", 1.0))
    db.commit()
    assert db.execute("SELECT COUNT(*) FROM %s" % t).fetchone() == (4,)
    with pytest.raises(g.GpudError):
        st.record_metrics([(1, "", "m", "", 0.0)])                # ErrEmptyComponentName
    db.close()


def _raw_ext_buffer(n_records=600, seed=3):
    """/dev/kmsg records whose messages are the line-matcher vectors, several per minute, some repeated inside one minute"""
    import numpy as np
    import synth
    rng = np.random.default_rng(seed)
    msgs = [l for l in synth.ext_lines() + synth.EXT_EDGE_LINES if l and "\n" not in l]
    recs, usec = [], 5_000_000
    for i in range(n_records):
        usec += int(rng.choice([1000, 200_000, 7_000_000, 45_000_000]))
        m = msgs[int(rng.integers(0, len(msgs)))] if rng.random() < 0.7 else "usb 1-%d: new high-speed USB device" % i
        recs.append("%d,%d,%d,-;%s" % (int(rng.integers(0, 8)), 1000 + i, usec, m))
        if rng.random() < 0.15:
            recs.append(recs[-1].replace(",%d,%d," % (1000 + i, usec), ",%d,%d," % (9000 + i, usec + 500)))   # same message, same minute
    return "\n".join(recs).encode()


def _oracle_hits(buf):
    from oracle import pyoracle as O
    want, n_units = O.scan_raw_kmsg(buf, ext=True)
    hits = []
    for w in want:
        if w["kind"] < 3:
            continue
        h = g.XidHit()
        h.unit_index, h.kind, h.kmsg_priority, h.kmsg_seq, h.kmsg_usec = w["line"], w["kind"], w["kmsg"][0], w["kmsg"][1], w["kmsg"][2]
        cap = w["capture"]
        assert len(cap) <= 39
        h.dev_len, h.device = len(cap), cap
        hits.append(h)
    return hits, want


@pytest.mark.parametrize("component", ["disk", "infiniband", "cpu", "os", "nccl", "peermem"])
def test_kmsg_syncer_rows_match_the_reference_flow(store, golden, component):
    from oracle import pyoracle as O
    st, path = store
    G = golden("store_sql.json")
    buf = _raw_ext_buffer()
    hits, want = _oracle_hits(buf)
    boot = 1_740_000_000
    # restatement of Syncer.sync (pkg/kmsg/syncer.go:73-143) over the records, in order
    recs = buf.split(b"\n")
    seen, rows = set(), []
    for rec in recs:
        try:
            _p, _s, usec, msg = O.parse_kmsg_line(0, rec.decode("latin-1"))
        except ValueError:
            continue
        name, message = O.component_match(component, msg.encode("latin-1"))
        if not name:
            continue
        t = boot + usec // 1_000_000
        key = (t - t % 60, name + "_" + message)
        if key in seen:
            continue
        seen.add(key)
        row = (t, name, "Warning", message, None)
        if row not in rows:
            rows.append(row)
    assert len(rows) >= 5
    sy = st.syncer(component)
    t = st.event_table(component)
    half = len(hits) // 2
    n1 = st.syncer_feed(sy, component, hits[:half], buf, boot, now_unix=boot + 100)
    n2 = st.syncer_feed(sy, component, hits[half:], buf, boot, now_unix=boot + 200)
    db = sqlite3.connect(path)
    got = list(db.execute(G["event_get"]["sql"].format(table=t), (0,)))
    assert n1 + n2 == len(got) == len(rows) and sorted(got) == sorted(rows)
    assert [r[0] for r in got] == sorted((r[0] for r in got), reverse=True)
    # a restarted syncer (cold dedup cache, e.g. after a gpud restart) re-reading the same kmsg inserts nothing: Find
    sy2 = st.syncer(component)
    assert st.syncer_feed(sy2, component, hits, buf, boot, now_unix=boot + 10_000) == 0
    db.close()


# ---- Bucket.Find / compareEvent (pkg/eventstore/database.go:277-324, 459-469) ------------------------------------------
def test_find_event_compares_extra_info_as_a_map(store, golden):
    st, path = store
    G = golden("eventstore_cases.json")
    t = st.event_table("test_table")
    js = lambda m: json.dumps(m, separators=(",", ":"), sort_keys=True) if m is not None else ""
    for i, r in enumerate(G["compare_event"]["rows"]):            # TestCompareEvent: stored event A, searched event B
        ts = 1_750_000_000 + i
        st.insert_event(t, ts, "kmsg", "Warning", "", js(r["a"]))
        assert st.find_event(t, ts, "kmsg", "Warning", "", js(r["b"])) == r["expected"], r["name"]
        # the text of the JSON does not matter, the map does: reversed key order and extra spaces
        spaced = "{ " + " , ".join("%s : %s" % (json.dumps(k), json.dumps(v)) for k, v in reversed(list(r["b"].items()))) + " }"
        assert st.find_event(t, ts, "kmsg", "Warning", "", spaced) == r["expected"], r["name"]
    # TestFindEvent (database_test.go:309): absent before the insert, present after
    ev = (1_750_001_000, "kmsg", "Warning", "", js({"a": "b"}))
    assert not st.find_event(t, *ev)
    st.insert_event(t, *ev)
    assert st.find_event(t, *ev)
    # TestFindEventPartialMatch (:351): same timestamp / name / type, different details -> nil
    assert not st.find_event(t, 1_750_001_000, "kmsg", "Warning", "", js({"a": "c"}))
    # TestFindEventMultipleMatches (:391): two rows of the same (timestamp, name, type); the one with the equal map is found
    st.insert_event(t, 1_750_002_000, "kmsg", "Warning", "", js({"a": "b", "c": "d"}))
    st.insert_event(t, 1_750_002_000, "kmsg", "Warning", "", js({"a": "b"}))
    assert st.find_event(t, 1_750_002_000, "kmsg", "Warning", "", js({"a": "b"}))
    assert st.find_event(t, 1_750_002_000, "kmsg", "Warning", "", js({"c": "d", "a": "b"}))
    assert not st.find_event(t, 1_750_002_000, "kmsg", "Warning", "", js({"c": "d"}))
    # message: part of the WHERE clause only when non-empty (database.go:291-298)
    st.insert_event(t, 1_750_003_000, "kmsg", "Warning", "hello", "")
    assert st.find_event(t, 1_750_003_000, "kmsg", "Warning", "", "") and st.find_event(t, 1_750_003_000, "kmsg", "Warning", "hello", "")
    assert not st.find_event(t, 1_750_003_000, "kmsg", "Warning", "other", "") and not st.find_event(t, 1_750_003_000, "kmsg", "Fatal", "", "")
    # escapes decode to the same strings (json.Unmarshal): \\u0041 is "A", a surrogate pair is one code point
    st.insert_event(t, 1_750_004_000, "kmsg", "Warning", "", json.dumps({"k": "A\n\U0001F600/"}, ensure_ascii=False))
    assert st.find_event(t, 1_750_004_000, "kmsg", "Warning", "", '{"k":"\\u0041\\n\\ud83d\\ude00\\/"}')
    assert not st.find_event(t, 1_750_004_000, "kmsg", "Warning", "", '{"k":"\\u0042\\n\\ud83d\\ude00\\/"}')


def test_stored_extra_info_is_read_like_unmarshal_if_valid(store, golden):
    st, path = store
    G = golden("eventstore_cases.json")
    t = st.event_table("test_table")
    db = sqlite3.connect(path)
    for i, r in enumerate(G["unmarshal_if_valid"]["rows"]):       # TestUnmarshalIfValid (database_test.go:1438)
        if r["name"] in ("valid JSON", "invalid JSON format"):     # rows about the test's own struct type (an int field), not map[string]string
            continue
        ts = 1_760_000_000 + i
        db.execute("INSERT INTO %s (timestamp, name, type, message, extra_info) VALUES (?, 'kmsg', 'Warning', NULL, ?)" % t, (ts, r["string"] if r["valid"] else None))
        db.commit()
        if r["expected_error"]:
            with pytest.raises(g.GpudError):
                st.find_event(t, ts, "kmsg", "Warning", "", "")
        else:
            assert st.find_event(t, ts, "kmsg", "Warning", "", "")          # NULL, "" and "null" all read back as "no ExtraInfo"
            assert st.find_event(t, ts, "kmsg", "Warning", "", "null") and st.find_event(t, ts, "kmsg", "Warning", "", "{}")
    # map[string]string: a non-string value is an Unmarshal error, a null value leaves ""
    db.execute("INSERT INTO %s (timestamp, name, type, message, extra_info) VALUES (1760000100, 'kmsg', 'Warning', NULL, '{\"key\":\"test\",\"value\":123}')" % t)
    db.execute("INSERT INTO %s (timestamp, name, type, message, extra_info) VALUES (1760000200, 'kmsg', 'Warning', NULL, '{\"key\":null}')" % t)
    db.commit()
    with pytest.raises(g.GpudError):
        st.find_event(t, 1760000100, "kmsg", "Warning", "", "")
    assert st.find_event(t, 1760000200, "kmsg", "Warning", "", '{"key":""}')
    db.close()


# ---- the Syncer's dedup options (pkg/kmsg/syncer.go:30-59, 84-155; syncer_test.go:62-383) -----------------------------
BASE = 1767270610                                   # time.Date(2026, 1, 1, 12, 30, 10, 0, time.UTC) of the reference's tests


def _syncer_case(st, path, G, bucket, cfg, sends, match):
    """the shape of every TestSyncer_* dedup test: a mock watcher channel, a matchFunc, options; returns the rows (time DESC)"""
    sy = st.syncer(bucket)
    st.syncer_configure(sy, **cfg)
    for dt, raw in sends:
        name, message = match(raw)
        if name:
            st.syncer_offer(sy, BASE + dt, name, message, now_unix=BASE + 1000)
    db = sqlite3.connect(path)
    rows = list(db.execute(G["event_get"]["sql"].format(table=st.event_table(bucket)), (0,)))
    db.close()
    return rows


def test_syncer_dedup_options_follow_the_references_tests(store, golden):
    import datetime
    assert BASE == int(datetime.datetime(2026, 1, 1, 12, 30, 10, tzinfo=datetime.timezone.utc).timestamp())
    st, path = store
    G = golden("store_sql.json")
    const = lambda raw: ("test_event", "constant parsed message") if raw else ("", "")
    pids = [(0, "raw message with pid 123"), (1, "raw message with pid 456"), (2, "raw message with pid 789")]
    # TestSyncer_Deduplication (:62): WithCacheKeyTruncateSeconds(60), three raw lines one parsed form -> exactly 1 event
    rows = _syncer_case(st, path, G, "test_dedup", dict(truncate_seconds=60), pids, const)
    assert len(rows) == 1 and rows[0][1:4] == ("test_event", "Warning", "constant parsed message")
    # TestSyncer_DisableDedup (:119): withDisableDedup -> 3 events
    assert len(_syncer_case(st, path, G, "test_disable_dedup", dict(disable_dedup=True), pids, const)) == 3
    # TestSyncer_EventDedupWindowFunc (:164): 5 min window for test_event; +0, +4 min, +6 min -> 2 events
    five = [("test_event", "", 300)]
    rows = _syncer_case(st, path, G, "test_event_dedup_window", dict(rules=five), [(0, "a"), (240, "b"), (360, "c")], const)
    assert len(rows) == 2
    # TestSyncer_EventDedupWindowFunc_BypassesGenericDedup (:216): generic 300 s, event window 1 min; +0, +2 min -> 2 events
    one = [("test_event", "", 60)]
    assert len(_syncer_case(st, path, G, "test_event_dedup_bypass_generic", dict(truncate_seconds=300, rules=one), [(0, "a"), (120, "b")], const)) == 2
    # TestSyncer_EventDedupWindowFunc_PreservesGenericDedupForOtherEvents (:266): base 12:30:00; generic 300 s keeps 1 generic, window 60 s keeps 2 custom
    def two(raw):
        return {"custom-1": ("custom_event", "custom parsed message"), "custom-2": ("custom_event", "custom parsed message"),
                "generic-1": ("generic_event", "generic parsed message"), "generic-2": ("generic_event", "generic parsed message")}.get(raw, ("", ""))
    rows = _syncer_case(st, path, G, "test_event_dedup_preserves_generic", dict(truncate_seconds=300, rules=[("custom_event", "", 60)]),
                        [(-10, "generic-1"), (110, "generic-2"), (-10, "custom-1"), (110, "custom-2"), (5, "unmatched")], two)
    assert len(rows) == 3 and sum(r[1] == "generic_event" for r in rows) == 1 and sum(r[1] == "custom_event" for r in rows) == 2
    # TestSyncer_DisableDedup_KeepsEventDedupWindowFunc (:334): disabled generic dedup, 5 min window still coalesces -> 1 event
    assert len(_syncer_case(st, path, G, "test_disable_dedup_keeps_event_window", dict(disable_dedup=True, rules=five), pids[:2], const)) == 1


def test_syncer_component_options_and_random_streams_match_the_model(store, golden):
    import numpy as np
    from oracle import pyoracle as O
    st, path = store
    G = golden("store_sql.json")
    E = golden("eventstore_cases.json")["infiniband_dedup_window"]
    assert E["constants"] == {"defaultKmsgEventDedupWindow": 300, "defaultAccessRegEventDedupWindow": 86400}
    for r in E["rows"]:                              # TestComponentKmsgEventDedupWindow_* (infiniband/component_test.go:261-315)
        assert O.infiniband_dedup_window(r["event"], r["message"]) == (r["window_seconds"], r["ok"]), r["name"]
    rng = np.random.default_rng(11)
    events = [(r["event"], r["message"]) for r in E["rows"]] + [("access_reg_failed", E["rows"][0]["message"].replace("d2", "3b")), ("port_module_high_temperature", "x")]
    configs = {"infiniband": (dict(truncate_seconds=300, window_func=O.infiniband_dedup_window), None),
               "peermem": (dict(truncate_seconds=300), None), "disk": (dict(truncate_seconds=300), None), "nccl": (dict(), None), "memory": (dict(), None),
               "custom": (dict(truncate_seconds=120, disable_dedup=True, window_func=lambda n, m: (90, True) if n == "other_event" else (0, n == "access_reg_failed")),
                          dict(truncate_seconds=120, disable_dedup=True, rules=[("other_event", "", 90), ("access_reg_failed", "", 0)])),
               "off": (dict(disable_dedup=True), dict(disable_dedup=True))}
    for i, (comp, (mcfg, pcfg)) in enumerate(configs.items()):
        model = O.KmsgSyncerModel(**mcfg)
        bucket = "model_%s" % comp
        sy = st.syncer(bucket)
        if pcfg is None:
            st.syncer_configure_component(sy, comp)
        else:
            st.syncer_configure(sy, **pcfg)
        t, now = 1_770_000_000, 1_770_000_000
        for _ in range(1500):
            t += int(rng.choice([0, 1, 20, 70, 400, 4000, 50_000]))
            now += int(rng.choice([0, 5, 100, 1000, 100_000]))
            name, msg = events[int(rng.integers(0, len(events)))]
            tt = t - int(rng.choice([0, 0, 30, 500]))            # kmsg timestamps are not always monotone across sources
            assert st.syncer_offer(sy, tt, name, msg, now) == model.offer(tt, name, msg, now), (comp, tt, name)
        db = sqlite3.connect(path)
        rows = list(db.execute(G["event_get"]["sql"].format(table=st.event_table(bucket)), (0,)))
        db.close()
        assert sorted((r[0], r[1], r[2], r[3]) for r in rows) == sorted(model.rows) and len(rows) > 10
    sy = st.syncer("model_bad")
    with pytest.raises(g.GpudError):
        st.syncer_configure_component(sy, "no-such-component")


def test_sxid_hits_are_persisted_like_the_component_does(store, golden):
    """sxid/component.go:433-469: error_sxid rows with an empty type, the decimal code as data, duplicates skipped; read back with the
    reference's query and resolved like resolveSXIDEvent"""
    from oracle import pyoracle as O
    st, path = store
    G = golden("store_sql.json")
    t = st.event_table("accelerator-nvidia-error-sxid")
    lines = [r["input"] for r in golden("sxid_kmsg.json")["match"]["rows"]]
    hits = []
    for i, ln in enumerate(lines):
        m = O.sxid_match(ln.encode())
        if m is None:
            continue
        h = g.XidHit()
        h.unit_index, h.kind, h.code, h.kmsg_usec = i, 2, m["sxid"], (100 + i) * 1_000_000
        dev = m["device"].encode()
        h.device, h.dev_len = dev, len(dev)
        hits.append(h)
    assert hits
    xh = g.XidHit()
    xh.kind, xh.code = 1, 79                                    # an xid hit in the same list is not this component's
    n1 = st.insert_sxid_hits(t, hits + [xh], boot_unix=1_740_000_000, raw_kmsg=True)
    n2 = st.insert_sxid_hits(t, hits, boot_unix=1_740_000_000, raw_kmsg=True)          # the same scan again: all duplicates
    db = sqlite3.connect(path)
    rows = list(db.execute(G["event_get"]["sql"].format(table=t), (0,)))
    db.close()
    uniq = {(1_740_000_000 + h.kmsg_usec // 1_000_000, h.code, bytes(h.device).rstrip(b"\0").decode()) for h in hits}
    assert n1 == len(rows) == len(uniq) and n2 == 0
    for ts, name, typ, msg, extra in rows:
        assert (name, typ, msg) == ("error_sxid", "", None)
        e = json.loads(extra)
        assert set(e) == {"data", "device_uuid"} and (ts, int(e["data"]), e["device_uuid"]) in uniq
        assert extra == json.dumps(e, separators=(",", ":"), sort_keys=True)       # json.Marshal of a map: sorted keys, compact
        r = O.resolve_sxid_event(typ, e["data"], e["device_uuid"])
        assert r is not None and r[1].startswith("SXID %s(" % e["data"])


def test_find_event_map_compare_property(store):
    """Bucket.Find's compareEvent is map equality, whatever the JSON text looks like: random string maps, either escaping style"""
    from hypothesis import given, settings, strategies as st_, HealthCheck
    st, path = store
    t = st.event_table("prop_table")
    text = st_.text(alphabet=st_.characters(blacklist_categories=("Cs",), blacklist_characters="\x00"), max_size=12)
    maps = st_.dictionaries(text, text, max_size=4)
    counter = [1_780_000_000]

    @settings(max_examples=150, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(maps, maps, st_.booleans(), st_.booleans())
    def prop(a, b, ascii_a, ascii_b):
        counter[0] += 1
        ts = counter[0]
        st.insert_event(t, ts, "kmsg", "Warning", "", json.dumps(a, ensure_ascii=ascii_a) if a else "")
        assert st.find_event(t, ts, "kmsg", "Warning", "", json.dumps(b, ensure_ascii=ascii_b, indent=1) if b else "") == (a == b)
        assert st.find_event(t, ts, "kmsg", "Warning", "", json.dumps(a, ensure_ascii=not ascii_a, sort_keys=True) if a else "null")
    prop()


def test_hw_slowdown_flow_through_the_event_store(store, golden):
    """hw-slowdown Check (component.go:262-407) on the library: clock-event readings -> HWSlowdownEvent -> Find / Insert -> read back with the
    reference's Get query -> frequency rule; event shape from TestCreateEventFromClockEvents (clock_events_test.go:337-382)"""
    import numpy as np
    from oracle import pyoracle as O
    st, path = store
    G = golden("store_sql.json")
    t = st.event_table("accelerator-nvidia-hw-slowdown")
    test_time = 1704067200                               # time.Date(2024, 1, 1, 0, 0, 0, 0, time.UTC)
    assert g.hw_slowdown_event_message(0, "GPU-123") == "" and O.hw_slowdown_event(0, "GPU-123", test_time) is None      # "no hardware slowdown reasons" -> nil
    assert not st.insert_hw_slowdown(t, test_time, 0x1 | 0x4, "GPU-123")                                                # idle / sw power cap are not hw slowdown
    for mask in (0x8, 0x40, 0x80, 0x8 | 0x40 | 0x80, 0x8 | 0x20 | 0x4, 0xFFFF):
        ev = O.hw_slowdown_event(mask, "GPU-123", test_time)
        assert g.hw_slowdown_event_message(mask, "GPU-123") == ev[3]
    rng = np.random.default_rng(21)
    now = 1_760_000_000
    model = []                                           # (time, message, uuid) rows the reference would hold
    for i in range(400):
        ts = now - int(rng.integers(0, 1500))
        uuid = "GPU-%d" % int(rng.integers(0, 3))
        mask = int(rng.choice([0, 0x1, 0x8, 0x40, 0x80, 0x48, 0x4]))
        ev = O.hw_slowdown_event(mask, uuid, ts)
        did = st.insert_hw_slowdown(t, ts, mask, uuid)
        if ev is None:
            assert not did
            continue
        key = (ts, ev[3], uuid)
        assert did == (key not in model)
        if did:
            model.append(key)
    db = sqlite3.connect(path)
    for window, thr in ((600, 0.6), (600, 0.1), (1200, 0.6), (60, 0.6), (0, 0.6)):
        rows = list(db.execute(G["event_get"]["sql"].format(table=t), (now - window,)))          # Bucket.Get(since)
        assert sorted((r[0], r[3], json.loads(r[4])["gpu_uuid"]) for r in rows) == sorted(k for k in model if k[0] > now - window)
        for r in rows:
            assert (r[1], r[2]) == ("hw_slowdown", "Warning") and json.loads(r[4]) == {"data_source": "nvml", "gpu_uuid": json.loads(r[4])["gpu_uuid"]}
        times = [r[0] for r in rows]
        want = O.hw_slowdown_check([k[0] for k in model], now, window, thr)                       # the rule applies the window itself
        got = g.hw_slowdown_check(times, now, window, thr)
        assert (["Healthy", "Degraded", "Unhealthy"][got[0]], got[2], got[3]) == (want[0], want[4], want[3]) and abs(got[1] - want[1]) < 1e-12
    db.close()


def test_event_statements_are_the_references(golden):
    """the text of every event statement the library prepares == the fmt.Sprintf result in pkg/eventstore/database.go (extracted)"""
    import ctypes as C
    G = golden("store_sql.json")
    L = g.lib()
    out = C.create_string_buffer(1024)
    t = "components_x_events_v0_5_0"
    for which, key, suffix in ((0, "event_insert", ""), (1, "event_find", ""), (2, "event_find", " AND message = ?"), (3, "event_get", ""), (4, "event_latest", ""), (5, "event_purge", "")):
        assert L.gpudh_store_event_sql(which, t.encode(), out, 1024) > 0
        assert out.value.decode() == G[key]["sql"].format(table=t) + suffix, key


def test_bucket_get_latest_purge(store, golden):
    """TestGetEventsTimeRange (:122), TestEmptyResults (:173), TestPurgePartial (:247), TestLatest (:1250) of pkg/eventstore/database_test.go"""
    st, path = store
    G = golden("store_sql.json")
    base = 1_765_000_000
    t = st.event_table("test_table")
    assert st.get_events(t, base - 3600) == [] and st.latest_event(t) is None and st.purge_events(t, base) == 0      # empty store
    for dt in (-600, -300, 0):
        st.insert_event(t, base + dt, "kmsg", "Warning")
    assert len(st.get_events(t, base - 900)) == 3 and len(st.get_events(t, base - 120)) == 1
    assert [r[0] for r in st.get_events(t, 0)] == [base, base - 300, base - 600]                   # newest first
    assert st.get_events(t, base) == []                                                           # strictly newer than since
    db = sqlite3.connect(path)
    assert [tuple(r) for r in db.execute(G["event_get"]["sql"].format(table=t), (base - 900,))] == [(r[0], r[1], r[2], r[3] or None, r[4] or None) for r in st.get_events(t, base - 900)]
    t2 = st.event_table("purge_table")
    st.insert_event(t2, base - 600, "kmsg", "Warning", "", '{"id":"old_event"}')
    st.insert_event(t2, base, "kmsg", "Warning", "", '{"id":"new_event"}')
    assert st.purge_events(t2, base - 300) == 1
    rows = st.get_events(t2, base - 900)
    assert len(rows) == 1 and rows[0][4] == '{"id":"new_event"}'
    assert not st.find_event(t2, base - 600, "test", "Warning", "", '{"id":"old_event"}')
    t3 = st.event_table("latest_table")
    for dt, typ, msg, eid in ((-10, "Warning", "old event", "event1"), (0, "Info", "latest event", "event2"), (-5, "Critical", "middle event", "event3")):
        st.insert_event(t3, base + dt, "test", typ, msg, json.dumps({"id": eid}, separators=(",", ":")))
    latest = st.latest_event(t3)
    assert latest == (base, "test", "Info", "latest event", '{"id":"event2"}')
    assert tuple(db.execute(G["event_latest"]["sql"].format(table=t3)).fetchone()) == latest
    assert st.purge_events(t3, base + 3600) == 3 and st.latest_event(t3) is None
    # a stored extra_info that does not unmarshal fails Get like scanRows does; capacity errors are reported, not truncated silently
    db.execute("INSERT INTO %s (timestamp, name, type, message, extra_info) VALUES (?, 'kmsg', 'Warning', NULL, '[1]')" % t, (base + 50,))
    db.commit()
    with pytest.raises(g.GpudError):
        st.get_events(t, base)
    db.execute("DELETE FROM %s WHERE timestamp = ?" % t, (base + 50,))
    db.commit()
    with pytest.raises(g.GpudError):
        st.get_events(t, 0, cap_rows=2)
    db.close()


def test_component_state_from_the_stores_matches_the_reference_flow(store, golden):
    """updateCurrentState (xid/component.go:581-611, sxid/component.go:478-507) with both buckets in SQLite: random histories of error events
    (JSON payloads as the library persists them, legacy decimal payloads), reboots in the os bucket, SetHealthy markers; the oracle side is
    trimEventsAfterSetHealthy + mergeEvents + evolveHealthyState over the same rows"""
    import numpy as np
    from oracle import pyoracle as O
    st, path = store
    rng = np.random.default_rng(17)
    devices = {"GPU-aaaa": "0000:04:00.0"}
    now = 1_766_000_000
    HEALTH = ["Healthy", "Degraded", "Unhealthy"]
    wire = O.ACTION_WIRE
    for trial in range(25):
        xt, ot, stt = st.event_table("xid-%d" % trial), st.event_table("os-%d" % trial), st.event_table("sxid-%d" % trial)
        xrows, orows, srows = [], [], []
        used = set()
        for _ in range(int(rng.integers(0, 14))):
            ts = now - int(rng.integers(1, 4 * 24 * 3600))                  # some fall outside the 3-day lookback
            while ts in used:
                ts -= 1                                                     # distinct timestamps: sort.Slice is not stable on ties
            used.add(ts)
            k = rng.random()
            if k < 0.25:
                name = str(rng.choice(["reboot", "reboot", "kernel_panic"]))   # the os bucket holds other events too; only reboots count
                st.insert_event(ot, ts, name, "Warning", "system reboot detected")
                orows.append((ts, name))
            elif k < 0.32:
                st.insert_event(xt, ts, "SetHealthy", "Info")
                xrows.append({"time": ts, "name": "SetHealthy"})
            elif k < 0.75:
                xid = int(rng.choice([13, 31, 63, 79, 94, 149, 99999]))
                typ = str(rng.choice(["Warning", "Critical", "Fatal"]))
                if rng.random() < 0.8:
                    act = str(rng.choice([wire[2], wire[3], wire[4]]))
                    data = '{"time":null,"data_source":"kmsg","device_uuid":"PCI:0000:04:00","xid":%d%s,"suggested_actions_by_gpud":{"description":"","repair_actions":["%s"]}}' % (
                        xid, ',"sub_code":4' if xid == 149 and rng.random() < 0.5 else "", act)
                else:
                    data = str(xid)
                st.insert_event(xt, ts, "error_xid", typ, "", json.dumps({"data": data, "device_uuid": "PCI:0000:04:00"}, separators=(",", ":"), sort_keys=True))
                xrows.append({"time": ts, "name": "error_xid", "type": typ, "data": data, "device_uuid": "PCI:0000:04:00"})
            else:
                code = int(rng.choice(list(O.SXID_DETAILS)[:30] + [99999]))
                st.insert_event(stt, ts, "error_sxid", "", "", json.dumps({"data": str(code), "device_uuid": "PCI:0000:05:00"}, separators=(",", ":"), sort_keys=True))
                srows.append({"time": ts, "name": "error_sxid", "type": "", "data": str(code), "device_uuid": "PCI:0000:05:00"})
        for lookback in (3 * 24 * 3600, 3600 * 12):
            since = now - lookback
            reboots = sorted(({"time": t, "name": "reboot"} for t, n in orows if n == "reboot" and t > since), key=lambda e: -e["time"])
            for rows, fn, is_sxid in ((xrows, st.xid_state, False), (srows, st.sxid_state, True)):
                local = sorted((e for e in rows if e["time"] > since), key=lambda e: -e["time"])
                local = O.trim_events_after_set_healthy(local)
                merged = sorted(reboots + local, key=lambda e: -e["time"])
                if is_sxid:
                    want = O.evolve_sxid_stored(merged)
                    got = fn(stt, ot, now, lookback)
                else:
                    want = O.evolve_healthy_state_stored(merged, devices, 2)
                    got = fn(xt, ot, now, lookback, 2, devices)
                assert (HEALTH[got[0]], got[1], got[2]) == (want["health"], (want["actions"] or [0])[0], want["reason"]), (trial, lookback, is_sxid, merged)
    # without a reboot store the component's own events still fold
    assert st.xid_state(st.event_table("xid-0"), None, now)[0] in (0, 1, 2)


def test_record_reboot_follows_the_references_rules(store):
    """recordEvent (pkg/host/event.go:85-132) with the sub-tests of TestRecordEvent (event_test.go:158-357): recent boot recorded, boot beyond
    the retention skipped, same boot again skipped, a boot 30 s after the stored one skipped, two minutes after it recorded, an older boot
    than the latest event skipped"""
    import datetime
    st, path = store
    utc = lambda *a: int(datetime.datetime(*a, tzinfo=datetime.timezone.utc).timestamp())
    t = st.event_table("os")
    now = utc(2025, 5, 21, 15, 0, 0)
    assert st.record_reboot(t, now, now - 3600)                                  # "recent reboot should record event"
    assert st.get_events(t, 0) == [(now - 3600, "reboot", "Warning", "system reboot detected 2025-05-21 14:00:00 +0000 UTC", "")]
    assert not st.record_reboot(t, now, now - 2 * 3 * 24 * 3600)                 # "old reboot should not record event"
    assert not st.record_reboot(t, now, now - 3 * 24 * 3600) and len(st.get_events(t, 0)) == 1     # exactly the retention: still skipped (>=)
    assert not st.record_reboot(t, now + 5, now - 3600)                          # "duplicate event should not be recorded"
    t2 = st.event_table("os-isolated")
    base = utc(2025, 1, 1, 12, 0, 0)
    st.insert_event(t2, base, "reboot", "Warning", "system reboot detected 2025-01-01 12:00:00 +0000 UTC")     # fmt.Sprintf("... %v", baseTime)
    assert not st.record_reboot(t2, base + 100, base)                            # found by Find (same time, name, type, message)
    assert not st.record_reboot(t2, base + 100, base + 30)                       # "less than a minute different should not be recorded"
    assert not st.record_reboot(t2, base + 100, base - 500)                      # the stored event is later than this boot
    assert st.record_reboot(t2, base + 200, base + 120)                          # "more than a minute different should be recorded"
    assert st.record_reboot(t2, base + 200, base + 180)                          # exactly a minute after the previous one: recorded (elapsed < time.Minute is false)
    assert [r[0] for r in st.get_events(t2, 0)] == [base + 180, base + 120, base]
    # the latest event of the bucket counts whatever its name (bucket.Latest, :116)
    st.insert_event(t2, base + 1000, "kernel_panic", "Warning", "x")
    assert not st.record_reboot(t2, base + 2000, base + 900) and st.record_reboot(t2, base + 2000, base + 1100)
