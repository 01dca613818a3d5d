"""Deterministic synthetic inputs shared by the tests, smoke() and bench.py (SURVEY.md §8d shapes)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0x67707564   # "gpud"


def golden(name):
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.load(f)


def gauge_stream(n_fields: int, n: int, seed: int = SEED, dtype=np.float64) -> np.ndarray:
    """[n, F] row-major polls.  Field k models a gauge base_k + A_k sin(2 pi t / P_k) + sigma_k N(0,1) with 0.1 % spikes;
    every 8th field is a monotone counter, every 8th+1 an integer-valued gauge (ties)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)[:, None]
    base = rng.uniform(30.0, 90.0, n_fields)[None, :]
    amp = rng.uniform(1.0, 10.0, n_fields)[None, :]
    per = rng.uniform(500.0, 50000.0, n_fields)[None, :]
    sig = rng.uniform(0.1, 2.0, n_fields)[None, :]
    x = base + amp * np.sin(2 * np.pi * t / per) + sig * rng.standard_normal((n, n_fields))
    spikes = rng.random((n, n_fields)) < 0.001
    x = np.where(spikes, x + 40.0, x)
    for k in range(0, n_fields, 8):
        x[:, k] = np.cumsum(rng.integers(0, 5, n)).astype(np.float64)
    for k in range(1, n_fields, 8):
        x[:, k] = np.round(x[:, k])
    return np.ascontiguousarray(x, dtype=dtype)


def thresholds_for(x: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(x.mean(axis=0) + 2.0 * x.std(axis=0))


def hit_lines():
    """Every Xid / SXid line the reference's tests exercise, plus the injectable messages for all catalog codes."""
    g = golden("xid_kmsg.json")
    s = golden("sxid_kmsg.json")
    lines = []
    for k in ("extract_xid", "extract_device", "match", "unknown_code"):
        lines += [r["input"] for r in g[k]["rows"]]
    for k in ("extended", "match_nvlink_examples"):
        lines += [r["logLine"] for r in g[k]["rows"]]
    lines += [r["line"] for r in g["nvlink_log_coverage"]["rows"]]
    lines += [r["line"] for r in g["short_match"]["rows"]]
    inj = g["inject_messages"]
    for c in range(1, 175):
        lines.append(inj["known"][str(c)]["message"] if str(c) in inj["known"] else inj["template"] % c)
    for c in (99999, 11111, 0):
        lines.append(inj["template"] % c)
    for k in ("extract_sxid", "extract_device", "match"):
        lines += [r["input"] for r in s[k]["rows"]]
    return lines


EDGE_LINES = [
    "NVRM: Xid (PCI:0000:05:00): 79, first NVRM: Xid (PCI:0000:06:00): 31, second",     # two anchors, leftmost wins
    "NVRM: Xid (invalid): x NVRM: Xid (PCI:0000:07:00): 13, later anchor is the match",
    "NVRM: Xid (PCI:0000:05:00) no code here: abc, then : 48, lazy finds the first ': digits,'",
    "NVRM: Xid (PCI:0000:05:00): 99999999999999999999, overflow then NVRM: GPU 0000:29:00.0: GPU has fallen off the bus.",
    "NVRM: Xid (PCI:0000:05:00): 0079, leading zeros",
    "NVRM: Xid (PCI:0000:05:00): 0, zero code NVRM: GPU 18:00.0: GPU has fallen off the bus",
    "NVRM: Xid (PCI:0000:04:00): 145, RLW_REMAP Nonfatal XC0 i0 Link 00 (0x1ffffffff 0x00000001 0x0 0x0)",  # intrinfo > 32 bit -> plain
    "NVRM: Xid (PCI:0000:04:00): 145, RLW_REMAP Nonfatal XC0 i0 Link 00 (0x00000004 0x00000001 0xfffffffff 0x5)",  # optional hex overflow skipped
    "NVRM: Xid (PCI:0000:04:00): 146, TLW_RX/TLW_RX_PIPE0 Fatal XC0 i0 Link 00 (0x0000000a 0x00000004 0x0 0x0)",   # digit in unit -> not extended
    "NVRM: Xid (PCI:0000:04:00): 149, NETIR_LINK_EVT/NETIR_LINK_DOWN Fatal XC1 i12 Link -1 (0x00a00011 0x00000000)",
    "NVRM: Xid (PCI:0000:04:00): 149, NETIR_LINK_EVT\tFatal\t\tXC0 i0 Link 08 (0x004505c6 0x00000000",               # truncated but 2 hex words
    "NVRM: Xid (PCI:0000:04:00): 200, pid=1, name=a, RLW_REMAP Nonfatal XC0 i0 Link 00 (0x00000004 0x00000001)",     # unknown base code
    "NVRM: Xid (PCI:0000:04:00): 145, pid=12, name=we,ird, RLW_REMAP Nonfatal XC0 i0 Link 00 (0x00000004 0x00000001)",
    "NVRM:   The NVIDIA GPU 0000:18:00.0 (PCI ID) has fallen off the bus and is not responding to commands.",
    "NVRM: GPU abcd:ef:01.0:   GPU has fallen off the bus",
    "NVRM: GPU 0000:29:00.1: GPU has fallen off the bus.",
    "xNVRM: Xid (0000:03:00): 14, glued prefix",
    "NVRM: Xid (PCI:): 14, empty device",
    "NVRM: Xid (PCI:00g0): 14, bad device",
    "nvidia-nvswitch0: SXid (PCI:0000:00:00.0): 20034, Fatal and NVRM: Xid (PCI:0000:01:00): 74, both on one line",
    "SXid nothing: here SXid (PCI:0000:a9:00.0): 12028, second SXid anchor has the code",
    "SXid (PCI:zz): 20034, device regex fails -> empty device",
    "SXid: 999999999999999999999999, overflow",
    "V X NV SX NVR NVRM SXi VRM: Xid",
    "",
    "NVRM: Xid (PCI:0000:05:00): 79",          # no trailing comma
]


def ext_lines():
    """every line of the nccl / peermem / infiniband / cpu / os / disk matcher tables of the reference (tests/golden/ext*_kmsg.json)"""
    G = golden("ext_kmsg.json")
    out = []
    for k in ("nccl_has", "nccl_match", "peermem_has", "peermem_match"):
        out += [r["line"] for r in G[k]["rows"]]
    G2 = golden("ext2_kmsg.json")
    for k, v in G2.items():
        out += [r["line"] for r in v["rows"] if "line" in r]
    return out


def ext_fuzz_lines(n: int, seed: int = SEED):
    """mutations of the matcher vectors: truncations, one-byte edits, duplicated / swapped halves, two vectors glued together
    (several anchors per line, patterns cut short, digits and colons moved) - the oracle decides what they mean"""
    rng = np.random.default_rng(seed)
    base = [l for l in ext_lines() + EXT_EDGE_LINES + PRIM_EDGE_LINES + [x for sq in stateful_sequences() for x in sq] if l]
    repl = list(" :,.[]()#!0179aWs-/") + ["", "  ", "::", "12", "task ", " seconds", "blocked for more than ", "nvme nvme", "in libnccl.so"]
    out = []
    for i in range(n):
        l = base[int(rng.integers(0, len(base)))]
        op = int(rng.integers(0, 6))
        if op == 0:
            l = l[: int(rng.integers(0, len(l) + 1))]
        elif op == 1:
            l = l[int(rng.integers(0, len(l))):]
        elif op == 2:
            p = int(rng.integers(0, len(l)))
            l = l[:p] + repl[int(rng.integers(0, len(repl)))] + l[p + int(rng.integers(0, 2)):]
        elif op == 3:
            p = int(rng.integers(0, len(l)))
            l = l[p:] + " " + l[:p]
        elif op == 4:
            m = base[int(rng.integers(0, len(base)))]
            l = l + (" " if rng.random() < 0.5 else "") + m
        else:
            p, q = sorted(int(x) for x in rng.integers(0, len(l) + 1, 2))
            l = l[:p] + l[q:]
        out.append(l.replace("\n", " "))
    return out


PEERMEM_LIT = "ERROR detected invalid context, skipping further processing"
EXT_EDGE_LINES = [
    "segfault atin libnccl.so",                                   # `.*` may be empty
    "segfault at in libnccl.s",                                   # literal cut short
    "segfault at 0 in libnccl_so",                                # `\.` is a literal dot
    "in libnccl.so then segfault at 0",                           # wrong order
    "in libnccl.so segfault at 0 ip 1 in libnccl.so.2",           # second library mention counts
    "segfault at 1 in libcuda.so segfault at 2 in libnccl.so",    # two anchors, one line
    "segfault a", "segf", "segfault", "xsegfault at y in libnccl.so.",
    "ERROR detected invalid context, skipping further processin",
    "ERRO", "ERROR detected", "ERROR detected invalid context, skipping further processing and more text",
    PEERMEM_LIT + " " + PEERMEM_LIT,
    "NVRM: Xid (PCI:0000:05:00): 79, pid=1, GPU has fallen off the bus. segfault at 0 in libnccl.so " + PEERMEM_LIT,
    "nvswitch0: SXid (PCI:0000:c1:00.0): 12028, Non-fatal, " + PEERMEM_LIT,
    "task a:123blocked for more than 5 seconds",                  # [\\d]+ gives a digit back to `.+`
    "task a:1blocked for more than 5 seconds",                    # ... and cannot here
    "task a:1 blocked for more than 5 seconds blocked for more than x seconds",
    "task a:12 blocked for more than x seconds blocked for more than 7 seconds",
    "task :1 blocked for more than 5 seconds", "task a: blocked for more than 5 seconds",
    "task a b task c:9 x blocked for more than 5 seconds", "INFO: task x:1 y blocked for more than 1 secondsblocked for more than 22 seconds",
    "soft lockup - CPU#1 stuck for 2s! [a:1]", "soft lockup - CPU# stuck for 2s! [a:1]", "soft lockup - CPU#1 stuck for 2s! [a:1x]",
    "soft lockup - CPU#1 stuck for 2s! [:1]", "soft lockup - CPU#12 stuck for 345s! [a]b:77]",
    "Detected insufficient power on the PCIe slot (27W)", "Detected insufficient power on the PCIe slot (27W", "Detected insufficient power on the PCIe slot ()",
    "Port module eventHigh Temperature", "High Temperature Port module event", "mlx5_cmd_out_errACCESS_REGfailed",
    "mlx5_cmd_out_err failed ACCESS_REG", "mlx5_cmd_out_err ACCESS_REG failed on 0000:d2:00.0x and 00000:d2:00.0 then 0000:d2:00.7.",
    "mlx5_cmd_out_err ACCESS_REG failed 0000:d2:00.8 abcd:ef:01.2", "x0000:d2:00.1 mlx5_cmd_out_err ACCESS_REG failed _0000:d2:00.1 -0000:d2:00.1",
    "VFS: file-max limit  reached", "VFS: file-max limit 5 reached", "VFS: file-max limit 5reached",
    "md/raid: Disk failure on  detected, failing array", "md/raid: Disk failure on detected, failing array", "md/raid1: Disk failure on x detected, failing arra",
    "block nvme: no available path - failing I/O", "block nvm: no available path - failing I/O",
    "nvme nvme1: I/O  timeout, reset controller", "nvme nvme1: I/O timeout, reset controller", "nvme nvme: I/O x timeout, reset controller",
    "nvme nvme12: Disabling device after reset failure", "nvme nvme12 : Disabling device after reset failure", "nvme nvme nvme3: Disabling device after reset failure",
    "Buffer I/O error on dev a,, logical block 1", "Buffer I/O error on dev ,, logical block 1", "Buffer I/O error on dev , logical block 1",
    "Buffer I/O error on dev dm-0, logical block x", "Buffer I/O error on dev dm-0 , logical block 1", "Buffer I/O error on dev a,b, logical block 9",
    "I/O error while writing superbloc", "attempt to access beyond end of devic", "Remounting filesystem read-onl",
    "Remounting filesystem read-only Remounting filesystem read-only", "I/O error while writing superblock and attempt to access beyond end of device",
    "segfault at" + " " * 300 + "in libnccl.so",                  # literals more than one chunk row apart
    "s" * 40 + "egfault at in libnccl.so",
]


def stateful_sequences():
    """every multi-line scenario of the os kernel-panic and memory OOM matcher tests (tests/golden/ext3_kmsg.json)"""
    G = golden("ext3_kmsg.json")
    seqs = [r["logLines"] for r in G["os.panic_detection"]["rows"]]
    seqs += [[st[0] for st in r["scenario"]] for r in G["os.panic_stateful"]["rows"]]
    seqs += [r["messages"] for r in G["memory.match_func"]["rows"] + G["memory.stream"]["rows"]]
    seqs += [[r["line"]] for k in ("os.panic_start", "os.panic_cpu_pid", "memory.container_name", "memory.process_pid", "memory.oom_start")
             for r in G[k]["rows"]]
    return [[l.replace("\n", " ") for l in s] for s in seqs]


PRIM_EDGE_LINES = [
    "Kernel panic", "Kernel Panic", "Kernel PANIC", "Kernel  panic", "kernel panic", "Kernel pani", "xKernel panicx",
    "CPU: 1 PID: 2 Comm: a", "CPU: 1 PID: 2 Comm: ", "CPU: 1 PID: 2 Comm:  a", "CPU:  1 PID: 2 Comm: a", "CPU: 1 PID: 2 Comm: a\tb",
    "CPU: 99999999999999999999 PID: 2 Comm: a", "CPU: 1 PID: 99999999999999999999 Comm: a", "CPU: 1 PID: 2 Comm: a CPU: 3 PID: 4 Comm: b",
    "CPU: x CPU: 5 PID: 6 Comm: c", "x invoked oom-killer:", "invoked oom-killer", "invoked oom-killer: invoked oom-killer:",
    "oom-kill:constraint=A,nodemask=B,cpuset=C,mems_allowed=D,oom_memcg=E,task_memcg=F,task=G,pid=7,uid=8",
    "oom-kill:constraint=A,nodemask=B,cpuset=C,mems_allowed=D,oom_memcg=E,task_memcg=F,task=G,pid=7,uid=8,pid=9,uid=10",
    "oom-kill:constraint=A,nodemask=B,nodemask=B2,cpuset=C,mems_allowed=D,oom_memcg=E,oom_memcg=E2,task_memcg=F,task=G,task=G2,pid=7,uid=8",
    "oom-kill:constraint=,nodemask=,cpuset=,mems_allowed=,oom_memcg=,task_memcg=,task=,pid=,uid=",
    "oom-kill:constraint=A,nodemask=B,cpuset=C,mems_allowed=D,oom_memcg=E,task_memcg=F,task=G,pid=x7,uid=8",
    "oom-kill:constraint=A,nodemask=B,cpuset=C,mems_allowed=D,oom_memcg=E,task_memcg=F,task=G,uid=8,pid=7",
    "oom-kill:constraint=A,uid=0,nodemask=B,cpuset=C,mems_allowed=D,oom_memcg=/a,task_memcg=/b,task=G,pid=12,uid=8 oom-kill:constraint=Z,nodemask=,cpuset=,mems_allowed=,oom_memcg=/c,task_memcg=/d,task=H,pid=13,uid=9",
    "Task in /a killed as a result of limit of /b", "Task in  killed as a result of limit of ", "Task in /a/../b//c/. killed as a result of limit of ../x",
    "Task in /a killed as a result of limit of /b killed as a result of limit of /c", "Task in /a killed as a result of limit of",
    "Killed process 12 (a)", "Killed process 12 ()", "Killed process 12 (a) (b) c)", "Killed process  12 (a)", "Killed process 12(a)",
    "Killed process 99999999999999999999 (a)", "Killed process 12 (a", "Killed process 0 (zero)",
]


def stateful_stream(n_lines: int, seed: int = SEED):
    """a long log: noise lines with the reference's panic / OOM scenarios, single primitive lines and edge lines dropped in
    at random distances (so sequences overlap, time out, restart and interleave)"""
    rng = np.random.default_rng(seed)
    seqs = stateful_sequences()
    singles = [l for l in PRIM_EDGE_LINES]
    out = []
    while len(out) < n_lines:
        r = rng.random()
        if r < 0.08:
            sq = seqs[int(rng.integers(0, len(seqs)))]
            out += sq[: int(rng.integers(1, len(sq) + 1))] if rng.random() < 0.3 else sq
        elif r < 0.2:
            out.append(singles[int(rng.integers(0, len(singles)))])
        else:
            out += ["[%d.%06d] usb 1-%d: noise line %d" % (len(out), i, i % 7, i) for i in range(int(rng.integers(1, 9)))]
    return out[:n_lines]


def ext_buffer(n_bytes: int, seed: int = SEED, hit_every: int = 50):
    """dmesg_buffer with the extra matchers' lines and decoys mixed in"""
    base = dmesg_buffer(n_bytes, seed, hit_every).split(b"\n")
    ext = [l.encode() for l in ext_lines() + EXT_EDGE_LINES + PRIM_EDGE_LINES]
    out = []
    for i, l in enumerate(base):
        out.append(l)
        if i % 37 == 5:
            out.append(ext[(i // 37) % len(ext)])
    return b"\n".join(out)


def dmesg_buffer(n_bytes: int, seed: int = SEED, hit_every: int = 1000):
    """~n_bytes of log text: the reference's fixture lines + noise + decoys, with one hit-line per ~hit_every lines."""
    rng = np.random.default_rng(seed)
    fix = [l for l in golden("xid_kmsg.json")["dmesg_xid_119"]["lines"] if l]
    kfix = []
    for rec in golden("pkg_kmsg.json")["fixture:kmsg.1.log"]["records"]:
        if ";" in rec:
            kfix.append(rec.split(";", 1)[1])
    hits = hit_lines() + EDGE_LINES
    decoys = ["Xid", "NVRM:", "SXid (invalid)", "NVRM: Xid", "VRM", "X", "V", "NV", "SX"]
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789 .:-_[]()=,/ABCDEFGHIJKLMNOPQRSTUVWXYZ", dtype=np.uint8)
    out, size, i = [], 0, 0
    while size < n_bytes:
        r = i % 10
        if i % hit_every == hit_every - 1:
            line = hits[(i // hit_every) % len(hits)]
        elif r < 3:
            line = kfix[(i // 10) % len(kfix)]
        elif r < 5:
            line = fix[(i // 10) % len(fix)]
        else:
            ln = int(rng.integers(40, 200))
            line = "[%12.6f] " % (i * 0.001) + alphabet[rng.integers(0, len(alphabet), ln)].tobytes().decode()
            if rng.random() < 0.05:
                line += " " + decoys[int(rng.integers(0, len(decoys)))]
        b = line.encode("utf-8", "surrogateescape") if isinstance(line, str) else line
        out.append(b)
        size += len(b) + 1
        i += 1
    return b"\n".join(out) + b"\n"


def raw_kmsg_buffer(n_records: int, seed: int = SEED):
    """/dev/kmsg style records 'prio,seq,usec,-;msg' with continuation lines and a few malformed records."""
    rng = np.random.default_rng(seed)
    hits = hit_lines() + EDGE_LINES
    recs = []
    for i in range(n_records):
        if i % 7 == 3:
            msg = hits[(i // 7) % len(hits)].replace("\n", "\\x0a")
        else:
            msg = "usb 1-%d: new high-speed USB device number %d using xhci_hcd" % (i % 9, i)
        meta = "%d,%d,%d,-" % (int(rng.integers(0, 8)), 1000 + i, 5000000 + 137 * i)
        if i % 53 == 17:
            meta = "x,%d,9,-" % i            # unparsable priority -> record skipped (watcher.go:161-165)
        if i % 59 == 23:
            rec = "no semicolon record %d" % i
        else:
            rec = meta + ";" + msg
        if i % 5 == 0:
            rec += "\n SUBSYSTEM=pci\n DEVICE=+pci:0000:%02x:00.0" % (i % 256)
        recs.append(rec)
    return ("\n".join(recs) + "\n").encode("utf-8", "surrogateescape")


def ib_series(n_series: int, seed: int = SEED, max_len: int = 200):
    """random (device, port) snapshot series: (ts, down, total_link_downed), time-ordered, with down runs of random length,
    link counters that sometimes move inside a run, equal timestamps, and lengths around the 32-wide step of the kernel"""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(n_series):
        n = int(rng.choice([0, 1, 2, 3, 31, 32, 33, 63, 64, 65, int(rng.integers(0, max_len))]))
        ts, t, tld, down = [], 1_700_000_000 + int(rng.integers(0, 10**6)), int(rng.integers(0, 50)), bool(rng.random() < 0.5)
        p_flip = float(rng.choice([0.02, 0.1, 0.3, 0.6]))
        for i in range(n):
            t += int(rng.choice([0, 1, 5, 10, 30, 60, 300]))
            if rng.random() < p_flip:
                down = not down
            if rng.random() < 0.05:
                tld += 1
            ts.append((t, down, tld))
        out.append(ts)
    return out
