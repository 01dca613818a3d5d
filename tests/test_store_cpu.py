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
