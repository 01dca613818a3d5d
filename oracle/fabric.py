"""CPU restatement of the whole-box NVLink / fabric verdict.  TEST INFRASTRUCTURE ONLY (see pyoracle.py header).

Follows  nvlink.Check aggregation          components/accelerator/nvidia/nvlink/component.go:164-311
         evaluateHealthStateWithThresholds  nvlink/evaluate_threshold.go:77-188 (+ component.go:375-397)
         FabricState.GetIssues              pkg/nvidia/nvml/device/fabric_state.go:115-177
         collectFabricState                 fabric-manager/fabric_state.go:67-113
Input: one dict per GPU (sorted-UUID order == gpu_index), the same quantities gpud_fabric_raw carries."""
from typing import Dict, List

P2P_UNPROBED = 0xFF
REASONS = ["NO_ISSUE", "P2P_FAILURE", "NO_ACTIVE_LINKS", "THRESHOLD_SATISFIED", "THRESHOLD_VIOLATED", "NO_THRESHOLD", "NO_DATA"]


def fabric_issue_bits(g: Dict) -> int:
    if not g.get("fabric_valid", 0):
        return 0
    b = 0
    if g["fabric_state"] != 3:           # nvml.GPU_FABRIC_STATE_COMPLETED
        b |= 0x01
    if g["fabric_status"] != 0:          # nvml.SUCCESS
        b |= 0x02
    if g["fabric_summary"] == 2:
        b |= 0x04
    if g["fabric_summary"] == 3:
        b |= 0x08
    m = g["fabric_health_mask"]
    for shift, bit in ((0, 0x10), (2, 0x20), (4, 0x40), (6, 0x80)):
        if (m >> shift) & 3 == 1:
            b |= bit
    return b


NVML_RETURN_NAMES = {0: "SUCCESS", 1: "ERROR_UNINITIALIZED", 2: "ERROR_INVALID_ARGUMENT", 3: "ERROR_NOT_SUPPORTED", 4: "ERROR_NO_PERMISSION",
                     6: "ERROR_NOT_FOUND", 9: "ERROR_DRIVER_NOT_LOADED", 10: "ERROR_TIMEOUT", 15: "ERROR_GPU_IS_LOST", 999: "ERROR_UNKNOWN"}


def get_issues(g: Dict) -> List[str]:
    """FabricState.GetIssues (pkg/nvidia/nvml/device/fabric_state.go:115-146) + getHealthMaskIssues (:150-177): sorted strings"""
    issues = []
    st = g["fabric_state"]
    if st != 3:
        issues.append("state=" + {0: "Not Supported", 1: "Not Started", 2: "In Progress", 3: "Completed"}.get(st, "Unknown(%d)" % st))
    if g["fabric_status"] != 0:
        issues.append("status=" + NVML_RETURN_NAMES.get(g["fabric_status"], "ERROR_%d" % g["fabric_status"]))
    if g["fabric_summary"] == 2:
        issues.append("summary=Unhealthy")
    elif g["fabric_summary"] == 3:
        issues.append("summary=Limited Capacity")
    issues += health_mask_issues(g["fabric_health_mask"])
    return sorted(issues)


def health_mask_issues(m: int) -> List[str]:
    out = []
    for shift, text in ((0, "bandwidth degraded"), (2, "route recovery in progress"), (4, "route unhealthy"), (6, "access timeout recovery in progress")):
        if (m >> shift) & 3 == 1:
            out.append(text)
    return out


def verdict(gpus: List[Dict], at_least: int) -> Dict:
    n = len(gpus)
    v = {"n_gpus": n, "active": 0, "inactive": 0, "unsupported": 0, "active_mask": 0, "inactive_mask": 0, "unsupported_mask": 0,
         "total_replay": 0, "total_recovery": 0, "total_crc": 0, "fabric_unhealthy_gpu_mask": 0, "fabric_issue_bits": [0] * 16}
    expected = False
    for g in gpus:
        i = g["gpu_index"]
        nl = g["n_links"]
        en = list(g["link_feature_enabled"][:nl])
        expected = expected or bool(g["system_expected_nvlink"])
        if not g["nvlink_supported"]:
            v["unsupported"] += 1; v["unsupported_mask"] |= 1 << i
        elif nl > 0 and all(en):               # len(States) > 0 && AllFeatureEnabled()   component.go:281
            v["active"] += 1; v["active_mask"] |= 1 << i
        else:
            v["inactive"] += 1; v["inactive_mask"] |= 1 << i
        v["total_replay"] += sum(g["link_replay_errors"][:nl])
        v["total_recovery"] += sum(g["link_recovery_errors"][:nl])
        v["total_crc"] += sum(g["link_crc_errors"][:nl])
        fb = fabric_issue_bits(g)
        v["fabric_issue_bits"][i] = fb
        if fb:
            v["fabric_unhealthy_gpu_mask"] |= 1 << i
    v["fabric_healthy"] = int(v["fabric_unhealthy_gpu_mask"] == 0)
    system_expected = n > 1 and expected                                   # component.go:184
    exp_pairs = probed = ok = 0
    ok_mask = obs = 0
    if n > 1:
        exp_pairs = n * (n - 1) // 2
        by_idx = {g["gpu_index"]: g for g in gpus}
        for i in sorted(by_idx):
            for j in sorted(by_idx):
                if i >= j:
                    continue
                st = by_idx[i]["p2p_status"][j]
                if st == P2P_UNPROBED:
                    continue
                probed += 1
                obs |= 1 << st
                if st == 0:
                    ok += 1
                    ok_mask |= (1 << i) | (1 << j)
    v.update(p2p_expected_pairs=exp_pairs, p2p_probed_pairs=probed, p2p_ok_pairs=ok, p2p_ok_gpu_mask=ok_mask,
             p2p_observed_status_mask=obs, required=at_least)
    health, reason = 0, 0
    p2p_failure = system_expected and n > 1 and probed > 0 and ok == 0         # component.go:389-397
    complete = exp_pairs != 0 and probed == exp_pairs                          # component.go:375-380
    if p2p_failure and complete:
        health, reason = 2, 1
    elif at_least <= 0:
        if system_expected and n > 0 and v["active"] == 0 and ok_mask == 0:
            health, reason = 2, 2
    elif n == 0:
        reason = 6
    elif v["active"] >= at_least:
        reason = 3
    else:
        health, reason = 2, 4
    v["nvlink_health"], v["nvlink_reason"] = health, reason
    return v


def suggest_reboot(v: Dict) -> bool:
    """setNVLinkSuggestedActions (nvlink/evaluate_threshold.go:37-52, only reached on the unhealthy paths) +
    peerNVLinkStatusesSuggestReboot (component.go:398-415): RebootSystem when a GPU has inactive links, or when every pair
    was probed, none is OK and some observed status is not one of the five "not supported" codes."""
    if v["nvlink_health"] != 2:
        return False
    complete = v["p2p_expected_pairs"] != 0 and v["p2p_probed_pairs"] == v["p2p_expected_pairs"]
    other = v["p2p_observed_status_mask"] & ~0b0111110
    return v["inactive"] > 0 or (complete and v["p2p_ok_pairs"] == 0 and other != 0)


P2P_CODES = ["OK", "CNS", "GNS", "TNS", "DR", "NS", "U"]                       # nvlink/p2p.go:11-19


def reason_string(v: Dict, uuids: List[str]) -> str:
    """cr.reason after evaluateHealthStateWithThresholds (nvlink/evaluate_threshold.go:11-35,77-188; component.go:307)"""
    def names(mask):
        return ",".join(uuids[i] if i < len(uuids) and uuids[i] else "GPU-%d" % i for i in range(16) if mask >> i & 1)

    def details(reason):
        parts = []
        if v["p2p_probed_pairs"] > 0 and v["p2p_ok_pairs"] == 0 and v["p2p_observed_status_mask"]:
            parts.append("peer nvlink p2p statuses=" + ",".join(sorted(P2P_CODES[c] for c in range(7) if v["p2p_observed_status_mask"] >> c & 1)))
        if v["inactive_mask"]:
            parts.append("inactive nvlinks=" + names(v["inactive_mask"]))
        if v["unsupported_mask"]:
            parts.append("unsupported nvlinks=" + names(v["unsupported_mask"]))
        return reason if not parts else "%s (%s)" % (reason, "; ".join(parts))
    r, n = v["nvlink_reason"], v["n_gpus"]
    if r == 0:
        return "all %d GPU(s) were checked, no nvlink issue found" % n
    if r == 1:
        return details("no GPU pairs report NVLink P2P connectivity on %d-GPU NVLink-capable system" % n)
    if r == 2:
        return details("no GPUs report active nvlink links on %d-GPU NVLink-capable system" % n)
    if r == 3:
        return "nvlink threshold satisfied: require >=%d GPUs with all links active; got %d" % (v["required"], v["active"])
    if r == 4:
        return details("nvlink threshold violated: require >=%d GPUs with all links active; got %d" % (v["required"], v["active"]))
    return {5: "nvlink threshold not set (skipped evaluation)", 6: "no nvlink data (skipped evaluation)"}[r]


def report_reason(gpus: List[Dict], uuids: List[str]):
    """collectFabricState (fabric-manager/fabric_state.go:67-113) -> (healthy, reason)"""
    reasons = []
    for g, u in zip(gpus, uuids):
        issues = get_issues(g) if g.get("fabric_valid", 0) else []
        if issues:
            reasons.append("GPU %s: %s" % (u, ", ".join(issues)))
    return (not reasons), "; ".join(sorted(reasons))
