"""Model check (numpy only) of k_scan_filter's byte-parallel pre-filter (gpud_b200/csrc/kmsg_scan.cu, DESIGN.md "Scanner in detail"):
the pre-filter must flag a SUPERSET of the lanes that own an anchor - a missed lane is a missed hit, a spurious one only costs the
exact re-test.  A lane is 16 bytes plus the next four (look-ahead word); the arithmetic below is the kernel's, on uint32 words:

  "Xid" trigram    wx = w ^ 'XXXX';  b = wx | (y1 ^ 0x31313131) | (y2 ^ 0x3c3c3c3c)  with y1 / y2 = the same words one / two bytes on;
                   flag = (b - 0x01010101) & ~b & 0x80808080 over the four words
  "fallen off the bus"   one of the lane's four ALIGNED words equals "fall", "alle", "llen" or "len "

Ownership: the lane whose 16 bytes hold the 'X' of an "Xid" owns "Xid " there and "SXid" one byte earlier; the lane that holds the first
4-aligned word inside an occurrence of "fallen off the bus" owns it."""
import numpy as np


def _lanes(buf: bytes):
    n = (len(buf) + 15) // 16 * 16
    a = np.zeros(n + 16, dtype=np.uint8)
    a[: len(buf)] = np.frombuffer(buf, dtype=np.uint8)
    w = a.view("<u4")                                      # word j = bytes 4j .. 4j+3
    nl = n // 16
    return w[: 4 * nl + 4], nl


def _funnel(lo, hi, sh):
    return ((lo >> np.uint32(sh)) | (hi << np.uint32(32 - sh))).astype(np.uint32)


def prefilter_flags(buf: bytes):
    w, nl = _lanes(buf)
    flag = np.zeros(nl, dtype=bool)
    one, top = np.uint32(0x01010101), np.uint32(0x80808080)
    fall = [np.uint32(v) for v in (0x6c6c6166, 0x656c6c61, 0x6e656c6c, 0x206e656c)]
    for k in range(4):
        wk, wn = w[k: 4 * nl: 4], w[k + 1: 4 * nl + 1: 4]            # word k of every lane and the word after it (word 4 = look-ahead)
        wx, wxn = wk ^ np.uint32(0x58585858), wn ^ np.uint32(0x58585858)
        b = wx | (_funnel(wx, wxn, 8) ^ np.uint32(0x31313131)) | (_funnel(wx, wxn, 16) ^ np.uint32(0x3c3c3c3c))
        flag |= (((b - one) & ~b) & top) != 0
        for f in fall:
            flag |= wk == f
    return flag


def owners(buf: bytes):
    own = set()
    i = buf.find(b"Xid")
    while i >= 0:
        own.add(i // 16)
        i = buf.find(b"Xid", i + 1)
    i = buf.find(b"fallen off the bus")
    while i >= 0:
        own.add(((i + 3) // 4 * 4) // 16)
        i = buf.find(b"fallen off the bus", i + 1)
    return own


def test_prefilter_flags_every_owner_lane():
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789 .:-_[]()=,/ABCDEFGHIJKLMNOPQRSTUVWXYZ\n", dtype=np.uint8)
    anchors = [b"NVRM: Xid (PCI:0000:05:00): 79, x", b"SXid (PCI:0000:05:00.0): 12028, y", b"GPU has fallen off the bus.", b"Xid", b"SXid", b"XXid", b"fallen off the bus"]
    n_own = 0
    for trial in range(300):
        n = int(rng.integers(40, 400))
        body = alphabet[rng.integers(0, len(alphabet), n)].tobytes()
        pos = sorted(rng.integers(0, n, int(rng.integers(1, 5))))
        out, last = [], 0
        for p in pos:
            out.append(body[last:p])
            out.append(anchors[int(rng.integers(0, len(anchors)))])
            last = p
        out.append(body[last:])
        buf = b"".join(out)
        flags = prefilter_flags(buf)
        for lane in owners(buf):
            n_own += 1
            assert flags[lane], (trial, lane, buf[max(0, lane * 16 - 4): lane * 16 + 24])
    assert n_own > 600


def test_every_alignment_and_the_look_ahead_word():
    for pad in range(0, 40):
        for lit in (b"Xid ", b"SXid", b"fallen off the bus"):
            buf = b"a" * pad + lit + b"zz"
            flags = prefilter_flags(buf)
            for lane in owners(buf):
                assert flags[lane], (pad, lit)


def test_prefilter_is_quiet_on_text_without_the_trigram():
    rng = np.random.default_rng(9)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789 .:-_[]()=,/ABCDEFGHIJKLMNOPQRSTUVWYZ\n", dtype=np.uint8)     # no 'X'
    buf = alphabet[rng.integers(0, len(alphabet), 1 << 16)].tobytes()
    flags = prefilter_flags(buf)
    own = owners(buf)
    assert flags.sum() <= len(own) + 8                  # only the four aligned words of "fallen ..." can flag here (chance hits are rare)


# ---- the separator arithmetic of k_scan_filter / k_scan_finish: exact counts, LINES and RAW_KMSG ---------------------------------
def _zero_bytes(x):
    m7 = np.uint32(0x7f7f7f7f)
    return ~(((x & m7) + m7) | x | m7)


def seps_in_block(block16: bytes, next_byte: int, raw_mode: bool, keep: int) -> int:
    """separators among the first `keep` bytes of a 16-byte block, the kernels' way (word arithmetic, look-ahead byte for RAW mode)"""
    w = np.frombuffer(block16 + bytes([next_byte, 0, 0, 0]), dtype="<u4")
    cnt = 0
    for k in range(4):
        z = _zero_bytes(w[k] ^ np.uint32(0x0a0a0a0a))
        if raw_mode:
            z &= ~_zero_bytes(_funnel(w[k: k + 1], w[k + 1: k + 2], 8)[0] ^ np.uint32(0x20202020))
        kk = keep - 4 * k
        if kk < 4:
            z = np.uint32(0) if kk <= 0 else z & np.uint32((1 << (8 * kk)) - 1)
        cnt += bin(int(z)).count("1")
    return cnt


def test_separator_arithmetic_is_exact():
    rng = np.random.default_rng(3)
    alphabet = np.frombuffer(b"ab \n\n\x0b\x0a \x09z", dtype=np.uint8)           # newline-heavy, with the bytes next to '\n' in value (VT, TAB)
    for trial in range(2000):
        blk = alphabet[rng.integers(0, len(alphabet), 17)].tobytes()
        keep = int(rng.integers(0, 17))
        for raw in (False, True):
            want = sum(1 for i in range(keep) if blk[i] == 0x0a and (not raw or blk[i + 1] != 0x20))
            assert seps_in_block(blk[:16], blk[16], raw, keep) == want, (trial, raw, keep, blk)
