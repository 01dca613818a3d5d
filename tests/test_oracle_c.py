"""Pin the C restatement (oracle/oracle.c + oracle/regex_bt.c: own backtracking engine over the reference's verbatim regex
strings) against the reference's golden vectors and against the independent Python `re` restatement."""
import numpy as np
import pytest

import synth
from oracle import coracle as CO
from oracle import pyoracle as O


def test_engine_compiles_all_six_patterns():
    assert CO.lib().orc_regex_ok() == 1


def _same(c, p):
    if p is None:
        assert c is None
        return
    assert c is not None and c.code == p.xid and c.device.decode() == p.device
    assert bool(c.extended) == (p.info is not None)
    if p.info is not None:
        assert (c.sub_code, c.intrinfo, c.error_status, c.link, c.unit.decode()) == \
               (p.info.sub_code, p.info.intrinfo, p.info.error_status, p.info.link, p.info.unit)
        assert bool(c.severity_fatal) == (p.info.severity == "Fatal")


def test_golden_match_vectors(golden):            # xid/kmsg_test.go:105-246, :289
    g = golden("xid_kmsg.json")
    for r in g["match"]["rows"] + g["unknown_code"]["rows"]:
        c = CO.xid_match(r["input"].encode())
        if r.get("expectNil"):
            assert c is None, r["name"]
        else:
            assert c is not None and c.code == r["expectedXid"] and c.device.decode() == r["expectedDevice"], r["name"]


def test_golden_extended_vectors(golden):         # xid/kmsg_extended_test.go:15-326
    for r in golden("xid_kmsg.json")["extended"]["rows"]:
        c = CO.xid_match(r["logLine"].encode())
        assert c is not None and c.extended == 1, r["name"]
        assert c.code == r["expectedXid"] and c.device.decode() == "PCI:" + r["expectedDeviceUUID"]
        assert (c.sub_code, c.unit.decode(), c.link, c.intrinfo, c.error_status) == \
               (r["expectedSubCode"], r["expectedSubCodeName"], r["expectedLink"], r["expectedIntrinfo"], r["expectedErrorStatus"])


def test_golden_sxid_vectors(golden):             # sxid/kmsg_test.go:102-170
    for r in golden("sxid_kmsg.json")["match"]["rows"]:
        c = CO.sxid_match(r["input"].encode())
        if r.get("expectNil"):
            assert c is None, r["name"]
        else:
            assert c.code == r["expectedSXid"] and c.device.decode() == r["expectedDevice"], r["name"]


def test_dmesg_fixture(golden):                   # xid/kmsg_test.go:248-287
    buf = "\n".join(golden("xid_kmsg.json")["dmesg_xid_119"]["lines"]).encode()
    hits, n_lines = CO.scan_lines(buf, threads=3)
    assert [(h.code, h.device.decode()) for h in hits] == [(119, "PCI:0000:9b:00")] * 5
    assert n_lines == buf.count(b"\n") + 1


def test_c_engine_equals_python_re_on_every_line():
    for line in synth.hit_lines() + synth.EDGE_LINES:
        b = line.encode()
        _same(CO.xid_match(b), O.xid_match(b))
        s = O.sxid_match(b)
        c = CO.sxid_match(b)
        assert (c is None) == (s is None)
        if s:
            assert (c.code, c.device.decode()) == (s["sxid"], s["device"])


def test_c_scan_equals_python_scan_on_synthetic_buffer():
    buf = synth.dmesg_buffer(1 << 20, hit_every=100)
    hits, n_lines = CO.scan_lines(buf, threads=4)
    want = O.scan_lines(buf)
    assert n_lines == buf.count(b"\n") + 1
    assert [(h.line, h.offset, h.kind, h.code, h.device.decode()) for h in hits] == \
           [(w["line"], w["offset"], w["kind"], w["code"], w["device"]) for w in want]
    assert len(want) > 80


def test_windows_c_equals_python():
    x = synth.gauge_stream(5, 2777, seed=21)
    thr = synth.thresholds_for(x)
    for W, a, qn, qd in ((1000, 0.0, 0, 0), (7, 0.3, 50, 100), (1024, 0.0, 999, 1000), (1, 0.0, 1, 1)):
        c = CO.windows_fields(np.ascontiguousarray(x.T), W, thr, a, qn, qd, threads=2)
        for f in range(x.shape[1]):
            w = O.window_aggregates(x[:, f], W, thr[f], a, qn or 99, qd or 100)
            for k in ("min", "max", "p99"):
                assert np.array_equal(c[k][f].view(np.uint64), w[k].view(np.uint64)), (W, k, f)
            assert np.array_equal(c["n_over"][f].astype(np.uint64), w["n_over"])
            assert np.allclose(c["mean"][f], w["mean"], rtol=1e-12, atol=0) and np.allclose(c["ema"][f], w["ema"], rtol=1e-9, atol=0)


def test_ext_matchers_golden_and_python(golden):   # nccl / peermem / infiniband / cpu / os / disk kmsg_matcher_test.go tables
    G = golden("ext_kmsg.json")
    for r in G["nccl_has"]["rows"]:
        assert bool(CO.ext_match(r["line"].encode())[0] & 1) == r["want"], r
    for r in G["peermem_has"]["rows"]:
        assert bool(CO.ext_match(r["line"].encode())[0] & 2) == r["want"], r
    lines = synth.ext_lines() + synth.EXT_EDGE_LINES + synth.PRIM_EDGE_LINES + synth.hit_lines()[:50] + synth.ext_fuzz_lines(3000)
    fired = set()
    for l in lines:
        b = l.encode()
        kinds = O.ext_match(b)
        fired.update(kinds)
        m, caps = CO.ext_match(b)
        assert m == sum(1 << (k - 3) for k in kinds), l
        for k in (8, 9):
            if k in kinds:
                assert caps[k] == O.ext_capture(k, b), l
        for k in (20, 22, 23, 24):                       # line primitives of the stateful matchers: every capture group
            if k in kinds:
                m = O.EXT_RE[k].search(b)
                assert CO.ext_groups(k, b)[: m.re.groups + 1] == [m.group(i) for i in range(m.re.groups + 1)], l
    assert fired == set(range(3, 25))


def test_c_ext_scan_equals_python_scan():
    buf = synth.ext_buffer(1_500_000, hit_every=40)
    ch, nl = CO.scan_lines(buf, ext=True)
    ph = O.scan_lines(buf, ext=True)
    assert nl == buf.count(b"\n") + 1
    assert [(h.line, h.kind, h.code, h.offset) for h in ch] == [(h["line"], h["kind"], h["code"], h["offset"]) for h in ph]
    assert len({h.kind for h in ch}) >= 20 and sum(h.kind >= 3 for h in ch) >= 80
    # the default scan is unchanged by the extra lines
    assert [(h.line, h.kind) for h in CO.scan_lines(buf)[0]] == [(h["line"], h["kind"]) for h in O.scan_lines(buf)]


def test_message_test_lines_agree_between_the_two_restatements(golden):
    """the kmsg lines of xid/health_state_test.go:318-690 (message / reason tests): C engine and Python `re` extract the same fields"""
    rows = golden("xid_messages.json")["from_lines"]["rows"]
    assert len(rows) == 15
    for r in rows:
        line = r["line"].encode()
        c, p = CO.xid_match(line), O.xid_match(line)
        _same(c, p)
        if "sub_code" in r:
            assert c.sub_code == r["sub_code"]
        if "xid" in r:
            assert c.code == r["xid"]
    # the header-only records of pkg/kmsg/watcher_test.go (negative priority / sequence, large timestamp) scan identically as raw kmsg
    calls = golden("pkg_kmsg.json")["parse_line_calls"]["rows"]
    recs = [r["input"].split(";", 1)[0] + ";NVRM: Xid (PCI:0000:05:00): 79, pid=1, GPU has fallen off the bus." for r in calls]
    buf = "\n".join(recs).encode()
    want, wn = O.scan_raw_kmsg(buf)
    assert wn == len(recs) == len(want)
