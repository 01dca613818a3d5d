"""CPU restatement of the InfiniBand port drop / flap scans.  TEST INFRASTRUCTURE ONLY (see pyoracle.py).

A series is the time-ordered snapshots of ONE (device, port): (ts, state, total_link_downed).  `down` below means the
reference's `snapshot.state != "active"`.  Timestamps and thresholds share one unit (the store keeps unix seconds).
Pinned by the reference's own tables (tests/golden/ib_scans.json <- infiniband/store/scan_drops_test.go, scan_flaps_test.go).
"""
import datetime


def find_drops(series, threshold):
    """devPortSnapshots.findDrops (components/accelerator/nvidia/infiniband/store/scan_drops.go:41-116).
    series: list of (ts, down, total_link_downed).  Returns None or {"down_since": ts, "index": i} (the latest snapshot)."""
    if len(series) <= 1:                                  # :43-45
        return None
    oldest = latest = None
    for i, (ts, down, tld) in enumerate(series):          # :49-68
        if not down:
            oldest = latest = None
            continue
        if oldest is None:
            oldest = i
        latest = i
    if oldest is None or latest is None:                  # :71-73
        return None
    if series[oldest][2] != series[latest][2]:            # :78-90
        return None
    if series[latest][0] - series[oldest][0] < threshold:  # :96-108
        return None
    return {"down_since": series[oldest][0], "index": latest}


def find_flaps(series, down_interval_threshold, flap_back_to_active_threshold):
    """devPortSnapshots.findFlaps (infiniband/store/scan_flaps.go:47-134).
    Returns None or {"down_since": ts of down1, "index": i of the revert that reached the threshold, "reverts": n}."""
    if len(series) < 3 or len(series) < flap_back_to_active_threshold:   # :50-52
        return None
    down1 = down2 = None
    reverts = []
    for i, (ts, down, _tld) in enumerate(series):         # :58-104
        if not down:
            if down1 is not None and down2 is not None:
                reverts.append((series[down1][0], i))
            down1 = down2 = None
            continue
        if down1 is None:
            down1 = i
        elif down2 is None:
            if ts - series[down1][0] < down_interval_threshold:
                continue
            down2 = i
    if len(reverts) < flap_back_to_active_threshold:      # :108-123
        return None
    since, idx = reverts[flap_back_to_active_threshold - 1]   # :133
    return {"down_since": since, "index": idx, "reverts": len(reverts)}


def rfc3339_utc(unix_seconds: int) -> str:
    return datetime.datetime.fromtimestamp(unix_seconds, datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ")


def drop_reason(device: str, port: int, down_since: int) -> str:          # scan_drops.go:112
    return "%s port %d down since %s" % (device, port, rfc3339_utc(down_since))


def flap_reason(device: str, port: int, down_since: int) -> str:          # scan_flaps.go:67
    return "%s port %d down since %s (and flapped back to active)" % (device, port, rfc3339_utc(down_since))
