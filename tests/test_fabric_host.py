"""CPU tests of the whole-box NVLink/fabric verdict oracle against the scenarios the reference's tests pin
(nvlink/evaluate_threshold_test.go:12-490, pkg/nvidia/nvml/device/fabric_state_test.go) — restated as inputs of the
per-GPU record form.  The GPU test (test_gpu_parity.py::test_fabric_pack_and_verdict) reuses SCENARIOS."""
from oracle import fabric as OF


def gpu(i, n, *, supported=1, expected=1, n_links=18, enabled=None, p2p=0, fabric=(1, 3, 1, 0, 0x00AA)):
    """fabric = (valid, state, summary, status, health_mask); mask 0xAA = every 2-bit field FALSE(2)"""
    en = [1] * 18 if enabled is None else list(enabled) + [0] * (18 - len(enabled))
    p = [0xFF] * 16
    for j in range(n):
        if j != i:
            p[j] = p2p if not isinstance(p2p, dict) else p2p.get((min(i, j), max(i, j)), 0)
    return {"gpu_index": i, "nvlink_supported": supported, "system_expected_nvlink": expected, "n_links": n_links,
            "link_feature_enabled": en, "link_replay_errors": [i + k for k in range(18)], "link_recovery_errors": [k % 3 for k in range(18)],
            "link_crc_errors": [(i * 7 + k) % 5 for k in range(18)], "p2p_status": p, "fabric_valid": fabric[0], "fabric_state": fabric[1],
            "fabric_summary": fabric[2], "fabric_status": fabric[3], "fabric_health_mask": fabric[4], "clique_id": 1}


def scenario(name, n=8):
    if name == "all_healthy":                       # Satisfied (evaluate_threshold_test.go:40)
        return [gpu(i, n) for i in range(n)], n
    if name == "gpu3_link7_down_threshold":         # PartialDegradationFailsWhenThresholdConfigured (:461)
        gs = [gpu(i, n) for i in range(n)]
        gs[3]["link_feature_enabled"][7] = 0
        return gs, n
    if name == "gpu3_link7_down_no_threshold":      # ImplicitFallbackDoesNotFailPartialDegradation (:430)
        gs = [gpu(i, n) for i in range(n)]
        gs[3]["link_feature_enabled"][7] = 0
        return gs, 0
    if name == "all_pairs_ns":                      # ImplicitFailureWhenPeerNVLinkP2PReportsNoOKPairs (:253)
        return [gpu(i, n, p2p=5) for i in range(n)], 0
    if name == "all_pairs_ns_with_threshold":       # ConfiguredThresholdStillFailsOnPeerNVLinkP2PFailure (:329)
        return [gpu(i, n, p2p=5) for i in range(n)], 1
    if name == "partial_probe_coverage":            # ImplicitPeerFailureSkippedWhenProbeCoverageIsPartial (:287)
        gs = [gpu(i, n, p2p=5) for i in range(n)]
        for j in range(2, n):
            gs[0]["p2p_status"][j] = 0xFF
        return gs, 0
    if name == "zero_active_no_threshold":          # ImplicitFailureWhenSystemExpectedNVLink (:229)
        return [gpu(i, n, enabled=[0] * 18, p2p=0xFF) for i in range(n)], 0
    if name == "zero_active_but_p2p_ok":            # ImplicitFallbackStaysHealthyWhenPeerP2POK (:397)
        return [gpu(i, n, enabled=[0] * 18, p2p=0) for i in range(n)], 0
    if name == "unsupported_violation":             # ViolationUnsupported (:89) / MixedInactiveAndUnsupported (:129)
        gs = [gpu(i, n) for i in range(n)]
        gs[1]["nvlink_supported"] = 0
        gs[2]["link_feature_enabled"] = [0] * 18
        return gs, n
    if name == "empty_states":                      # EmptyStates (:109): supported, zero links discovered -> inactive
        gs = [gpu(i, n) for i in range(n)]
        gs[0]["n_links"] = 0
        return gs, n
    if name == "fabric_summary_unhealthy":          # FailureInjector.GPUUUIDsWithFabricStateHealthSummaryUnhealthy (registry.go:58)
        gs = [gpu(i, n) for i in range(n)]
        gs[5] = gpu(5, n, fabric=(1, 3, 2, 0, 0x00AA))
        return gs, 0
    if name == "fabric_mask_and_state":             # GetIssues: state IN_PROGRESS, status error, bw degraded + route unhealthy
        gs = [gpu(i, n) for i in range(n)]
        gs[2] = gpu(2, n, fabric=(1, 2, 3, 9, (1 << 0) | (2 << 2) | (1 << 4) | (2 << 6)))
        gs[6] = gpu(6, n, fabric=(0, 0, 0, 0, 0))
        return gs, 0
    if name == "not_expected_single_domain":        # system without fabric support: nothing is implied
        return [gpu(i, n, expected=0, enabled=[0] * 18, p2p=5) for i in range(n)], 0
    raise KeyError(name)


SCENARIOS = ["all_healthy", "gpu3_link7_down_threshold", "gpu3_link7_down_no_threshold", "all_pairs_ns", "all_pairs_ns_with_threshold",
             "partial_probe_coverage", "zero_active_no_threshold", "zero_active_but_p2p_ok", "unsupported_violation", "empty_states",
             "fabric_summary_unhealthy", "fabric_mask_and_state", "not_expected_single_domain"]

# (health, reason) the Go tests assert for each scenario: 0 Healthy / 2 Unhealthy ; reason ids as GPUD_NVLINK_*
EXPECT = {"all_healthy": (0, 3), "gpu3_link7_down_threshold": (2, 4), "gpu3_link7_down_no_threshold": (0, 0), "all_pairs_ns": (2, 1),
          "all_pairs_ns_with_threshold": (2, 1), "partial_probe_coverage": (0, 0), "zero_active_no_threshold": (2, 2),
          "zero_active_but_p2p_ok": (0, 0), "unsupported_violation": (2, 4), "empty_states": (2, 4), "fabric_summary_unhealthy": (0, 0),
          "fabric_mask_and_state": (0, 0), "not_expected_single_domain": (0, 0)}


def test_nvlink_verdicts():
    for name in SCENARIOS:
        gpus, at_least = scenario(name)
        v = OF.verdict(gpus, at_least)
        assert (v["nvlink_health"], v["nvlink_reason"]) == EXPECT[name], name


def test_counts_and_masks():
    gpus, at_least = scenario("unsupported_violation")
    v = OF.verdict(gpus, at_least)
    assert (v["active"], v["inactive"], v["unsupported"]) == (6, 1, 1)
    assert v["unsupported_mask"] == 0b10 and v["inactive_mask"] == 0b100
    assert v["p2p_expected_pairs"] == 28 == v["p2p_probed_pairs"] == v["p2p_ok_pairs"]
    assert v["total_replay"] == sum(sum(g["link_replay_errors"]) for g in gpus)


def test_fabric_issues():
    # GetIssues (pkg/nvidia/nvml/device/fabric_state.go:115-177)
    gpus, _ = scenario("fabric_summary_unhealthy")
    v = OF.verdict(gpus, 0)
    assert v["fabric_healthy"] == 0 and v["fabric_unhealthy_gpu_mask"] == 1 << 5 and v["fabric_issue_bits"][5] == 0x04
    gpus, _ = scenario("fabric_mask_and_state")
    v = OF.verdict(gpus, 0)
    assert v["fabric_issue_bits"][2] == 0x01 | 0x02 | 0x08 | 0x10 | 0x40
    assert v["fabric_issue_bits"][6] == 0 and v["fabric_unhealthy_gpu_mask"] == 1 << 2
    gpus, _ = scenario("all_healthy")
    assert OF.verdict(gpus, 8)["fabric_healthy"] == 1


def test_single_gpu_is_left_alone():
    # 1-GPU hosts never expect NVLink (component.go:166-184)
    g = [gpu(0, 1, enabled=[0] * 18)]
    v = OF.verdict(g, 0)
    assert (v["nvlink_health"], v["p2p_expected_pairs"]) == (0, 0)


def test_get_issues_tables_of_the_reference(golden):
    """FabricState.GetIssues / getHealthMaskIssues: the reference's own tables (fabric_state_test.go:10,87) against the oracle, the
    oracle's issue bits, and the library's host rendering"""
    import ctypes as C
    import gpud_b200 as g
    G = golden("fabric_issues.json")
    L = g.lib()
    out = C.create_string_buffer(512)
    for r in G["get_issues"]["rows"]:
        d = {"fabric_valid": 1, "fabric_state": r["state"], "fabric_status": r["status"], "fabric_summary": r["summary"], "fabric_health_mask": r["health_mask"]}
        assert OF.get_issues(d) == r["expected"], r["name"]
        assert bool(OF.fabric_issue_bits(d)) == bool(r["expected"]), r["name"]
        raw = g.FabricRaw()
        raw.fabric_valid, raw.fabric_state, raw.fabric_status, raw.fabric_summary, raw.fabric_health_mask = 1, r["state"], r["status"], r["summary"], r["health_mask"]
        n = L.gpud_fabric_issues(C.byref(raw), out, 512)
        assert n >= 0 and out.value.decode() == ", ".join(r["expected"]), r["name"]
    for r in G["health_mask_issues"]["rows"]:
        assert OF.health_mask_issues(r["mask"]) == r["expected"], r["name"]
        raw = g.FabricRaw()
        raw.fabric_valid, raw.fabric_state, raw.fabric_summary, raw.fabric_health_mask = 1, 3, 1, r["mask"]
        L.gpud_fabric_issues(C.byref(raw), out, 512)
        assert out.value.decode() == ", ".join(sorted(r["expected"])), r["name"]


def golden_threshold_case(r):
    """one TestEvaluateThresholds_* vector (tests/golden/nvlink_thresholds.json) as the per-GPU records the library takes, plus the
    (health, reason id, reboot) its assertions mean"""
    uu = [g["uuid"] for g in r["nvlinks"]]
    n = len(uu)
    gs = []
    for i, e in enumerate(r["nvlinks"]):
        if e["uuid"] in r["active"]:
            cls = "a"
        elif e["uuid"] in r["inactive"]:
            cls = "i"
        elif e["uuid"] in r["unsupported"]:
            cls = "u"
        else:                                       # the test left the lists empty: classify like nvlink.Check does (component.go:271-289)
            cls = "u" if not e["supported"] else ("a" if e["states"] and all(e["states"]) else "i")
        nl = len(e["states"])
        en = [1] * max(nl, 1) if cls == "a" else ([int(x) for x in e["states"]] if nl and not all(e["states"]) else [0] * nl)
        if cls == "a":
            nl = max(nl, 1)
        d = gpu(i, max(n, 1), supported=0 if cls == "u" else 1, expected=1 if r["system_expected"] else 0, n_links=nl if cls != "u" else 0,
                enabled=(en + [0] * 18)[:18], p2p=0xFF)
        gs.append(d)
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    bad = next((c for c in r["p2p_observed"] if c != 0), 5)
    for k, (i, j) in enumerate(pairs[: r["p2p_probed"]]):
        st = 0 if k < r["p2p_ok"] else bad
        gs[i]["p2p_status"][j] = st
        gs[j]["p2p_status"][i] = st
    assert r["p2p_expected"] in (0, len(pairs)), r["name"]
    want_h = {"Healthy": 0, "Unhealthy": 2, "": 0}[r["want_health"] or r["preset_health"]]
    txt = " ".join(r["want_reason_contains"]) + " " + r["want_reason_equal"]
    if "no GPU pairs report NVLink P2P connectivity" in txt:
        rid = 1
    elif "no GPUs report active nvlink links" in txt:
        rid = 2
    elif "reasonNoNVLinkData" in txt:
        rid = 6
    elif "satisfied" in txt:
        rid = 3
    elif want_h == 2 and r["at_least"] > 0:
        rid = 4
    elif want_h == 2:
        rid = 1 if r["p2p_probed"] else 2
    else:
        rid = 3 if (r["at_least"] > 0 and n) else 0
    return gs, r["at_least"], want_h, rid, r["want_reboot"]


def test_threshold_vectors_of_the_reference(golden):
    """every TestEvaluateThresholds_* function of nvlink/evaluate_threshold_test.go, its checkResult turned into per-GPU records"""
    rows = golden("nvlink_thresholds.json")["evaluate"]["rows"]
    assert len(rows) == 17
    for r in rows:
        gs, at_least, want_h, rid, reboot = golden_threshold_case(r)
        v = OF.verdict(gs, at_least)
        assert (v["nvlink_health"], v["nvlink_reason"]) == (want_h, rid), (r["name"], v["nvlink_health"], v["nvlink_reason"], want_h, rid)
        if reboot is not None:
            assert OF.suggest_reboot(v) == reboot, r["name"]
        import ctypes as C
        import gpud_b200 as g
        fv = g.FabricVerdict()
        for k, val in v.items():
            if k != "fabric_issue_bits":
                setattr(fv, k, val)
        assert bool(g.lib().gpud_fabric_suggest_reboot(C.byref(fv))) == OF.suggest_reboot(v), r["name"]
        if len(r["active"]) + len(r["inactive"]) + len(r["unsupported"]) == len(r["nvlinks"]) > 0:      # the test listed every GPU
            assert (v["active"], v["inactive"], v["unsupported"]) == (len(r["active"]), len(r["inactive"]), len(r["unsupported"])), r["name"]
        # the reason text: the fragments / exact strings the reference's test asserts (constants resolved from evaluate_threshold.go:11-14)
        uu = [e["uuid"] for e in r["nvlinks"]]
        text = OF.reason_string(v, uu)
        assert g.fabric_reason(fv, uu) == text, r["name"]
        for frag in r["want_reason_contains"]:
            assert frag in text, (r["name"], frag, text)
        eq = {"reasonNoNVLinkData": "no nvlink data (skipped evaluation)", "reasonNoThresholdConfigured": "nvlink threshold not set (skipped evaluation)"}.get(r["want_reason_equal"], r["want_reason_equal"])
        if eq and eq != "existing reason":               # "existing reason": the test pre-set cr.reason to a placeholder that evaluate leaves alone
            assert text == eq, (r["name"], text)


def test_fabric_reason_strings_on_the_scenarios():
    import gpud_b200 as g
    for name in SCENARIOS:
        for n in (8,):
            gpus, at_least = scenario(name, n)
            v = OF.verdict(gpus, min(at_least, n))
            fv = g.FabricVerdict()
            for k, val in v.items():
                if k != "fabric_issue_bits":
                    setattr(fv, k, val)
            uu = ["GPU-%08x" % (i * 2654435761 % 2 ** 32) for i in range(n)]
            assert g.fabric_reason(fv, uu) == OF.reason_string(v, uu), (name, n)
            assert g.fabric_reason(fv) == OF.reason_string(v, [])          # unnamed GPUs render as GPU-<index>


def test_nvlink_states_helpers_of_the_reference():
    """NVLinkStates.AllFeatureEnabled / Total*Errors (nvlink/nvlink.go:34-68) as the verdict sees them: the tables of
    nvlink_test.go:57-93 (two links on / one off / no states) and :95-127 (10+15, 20+25, 30+35)"""
    def one(enabled, replay=(), recovery=(), crc=()):
        d = gpu(0, 1, n_links=len(enabled), enabled=(list(enabled) + [0] * 18)[:18], p2p=0xFF)
        d["link_replay_errors"] = (list(replay) + [0] * 18)[:18]
        d["link_recovery_errors"] = (list(recovery) + [0] * 18)[:18]
        d["link_crc_errors"] = (list(crc) + [0] * 18)[:18]
        return OF.verdict([d], 0)
    v = one([1, 1])
    assert (v["active"], v["inactive"]) == (1, 0)                      # "All links enabled" -> AllFeatureEnabled true, States non-empty
    v = one([1, 0])
    assert (v["active"], v["inactive"]) == (0, 1)                      # "Some links disabled"
    v = one([])
    assert (v["active"], v["inactive"]) == (0, 1)                      # "Empty states": AllFeatureEnabled() is true but len(States) == 0 (component.go:271-289)
    v = one([1, 1], (10, 15), (20, 25), (30, 35))
    assert (v["total_replay"], v["total_recovery"], v["total_crc"]) == (25, 45, 65)


def test_fabric_report_reason_of_collect_fabric_state():
    """TestCollectFabricState_SortsEntriesAndReasons / _SortsReasonsAcrossMultipleGPUs (fabric-manager/fabric_state_test.go:609-744): the
    FabricState literals of those tests (nvml.h values: state COMPLETED 3 / IN_PROGRESS 2, status SUCCESS 0 / ERROR_UNKNOWN 999, summary
    HEALTHY 1 / UNHEALTHY 2 / LIMITED_CAPACITY 3, mask fields of two bits at shifts 0, 2, 4, 6 with TRUE = 1), given out of UUID order"""
    import gpud_b200 as g
    def raw(state, status, mask, summary):
        r = g.FabricRaw()
        r.fabric_valid, r.fabric_state, r.fabric_status, r.fabric_health_mask, r.fabric_summary = 1, state, status, mask, summary
        return r
    def d(r):
        return {"fabric_valid": 1, "fabric_state": r.fabric_state, "fabric_status": r.fabric_status, "fabric_health_mask": r.fabric_health_mask, "fabric_summary": r.fabric_summary}
    info_a = raw(3, 0, 0, 1)
    info_b = raw(2, 999, (1 << 0) | (1 << 4), 2)
    healthy, reason = g.fabric_report_reason([info_b, info_a], ["GPU-B", "GPU-A"])
    assert not healthy and reason == "GPU GPU-B: " + ", ".join(OF.get_issues(d(info_b)))
    assert reason == "GPU GPU-B: bandwidth degraded, route unhealthy, state=In Progress, status=ERROR_UNKNOWN, summary=Unhealthy"
    state_a = raw(3, 0, 1 << 2, 3)                      # route recovery in progress + limited capacity
    info_c = raw(3, 0, 0, 0)
    healthy, reason = g.fabric_report_reason([info_c, info_b, state_a], ["GPU-C", "GPU-B", "GPU-A"])
    want_a = "GPU GPU-A: " + ", ".join(OF.get_issues(d(state_a)))
    want_b = "GPU GPU-B: " + ", ".join(OF.get_issues(d(info_b)))
    assert not healthy and reason == want_a + "; " + want_b == OF.report_reason([d(info_c), d(info_b), d(state_a)], ["GPU-C", "GPU-B", "GPU-A"])[1]
    assert g.fabric_report_reason([info_a, info_c], ["GPU-A", "GPU-C"]) == (True, "") and g.fabric_report_reason([], []) == (True, "")
    for name in SCENARIOS:                              # and on every scenario of the synthetic boxes
        gpus, _ = scenario(name, 8)
        raws = []
        for x in gpus:
            r = raw(x["fabric_state"], x["fabric_status"], x["fabric_health_mask"], x["fabric_summary"])
            r.fabric_valid = x["fabric_valid"]
            raws.append(r)
        uu = ["GPU-%02d" % (7 - i) for i in range(8)]
        assert g.fabric_report_reason(raws, uu) == OF.report_reason(gpus, uu), name


def test_fabric_status_text_follows_the_installed_nvml_error_string():
    """status.Error() is go-nvml's constant name until libnvidia-ml is loaded and nvmlErrorString's text afterwards; the library takes the
    function the same way (gpud_set_nvml_error_string), default = the constant names the reference's unit tests see"""
    import ctypes as C
    import gpud_b200 as g
    L = g.lib()
    r = g.FabricRaw()
    r.fabric_valid, r.fabric_state, r.fabric_status, r.fabric_summary = 1, 3, 999, 1
    out = C.create_string_buffer(256)
    L.gpud_fabric_issues(C.byref(r), out, 256)
    assert out.value.decode() == "status=ERROR_UNKNOWN"
    texts = {999: b"Unknown Error", 3: b"Not Supported", 15: b"GPU is lost"}
    bufs = {k: C.create_string_buffer(v) for k, v in texts.items()}          # the callback hands out pointers that stay valid, like nvmlErrorString
    other = C.create_string_buffer(b"?")
    FN = C.CFUNCTYPE(C.c_void_p, C.c_int32)
    cb = FN(lambda code: C.addressof(bufs.get(code, other)))
    try:
        L.gpud_set_nvml_error_string(cb)
        for code, text in texts.items():
            r.fabric_status = code
            L.gpud_fabric_issues(C.byref(r), out, 256)
            assert out.value.decode() == "status=" + text.decode()
        assert g.fabric_report_reason([r], ["GPU-A"]) == (False, "GPU GPU-A: status=GPU is lost")
    finally:
        L.gpud_set_nvml_error_string(None)
    r.fabric_status = 999
    L.gpud_fabric_issues(C.byref(r), out, 256)
    assert out.value.decode() == "status=ERROR_UNKNOWN"
    import torch
    if not torch.cuda.is_available():
        assert L.gpud_nvml_error_strings_from_driver() != 0        # no driver library on this host: refused, nothing installed
