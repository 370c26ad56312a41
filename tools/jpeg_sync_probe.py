"""Design probe for a GPU Huffman stage: how quickly does a baseline-JPEG entropy decoder
that starts at an ARBITRARY bit offset lock onto the true decode (same bit position, same coefficient index, same block
of the MCU)?  Self-synchronisation is what a parallel decoder without restart markers relies on: every thread starts at
the beginning of its sub-sequence in a guessed state, overruns into the next one, and stops where its state meets the
state the next thread reached.  The shorter the lock-in distance, the shorter the sub-sequences can be.

Pure Python on the de-stuffed scan of one interleaved JPEG (encoded here with cv2); prints one JSON line.
    python tools/jpeg_sync_probe.py [--quality 75] [--starts 400]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def parse(data):
    """-> (dc tables, ac tables, component list [(h, v, td, ta)], de-stuffed scan bits)"""
    i, dc, ac, comps, frame = 2, {}, {}, [], {}
    while i < len(data):
        assert data[i] == 0xFF
        m = data[i + 1]
        i += 2
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        seg = (data[i] << 8) | data[i + 1]
        s = data[i + 2:i + seg]
        if m == 0xC4:
            k = 0
            while k < len(s):
                tc, th = s[k] >> 4, s[k] & 15
                counts = s[k + 1:k + 17]
                nv = sum(counts)
                vals = s[k + 17:k + 17 + nv]
                table, code, v = {}, 0, 0
                for ln in range(1, 17):
                    for _ in range(counts[ln - 1]):
                        table[(ln, code)] = vals[v]
                        code += 1
                        v += 1
                    code <<= 1
                (ac if tc else dc)[th] = table
                k += 17 + nv
        elif m == 0xC0:
            for c in range(s[5]):
                frame[s[6 + 3 * c]] = (s[7 + 3 * c] >> 4, s[7 + 3 * c] & 15)
        elif m == 0xDD:
            assert ((s[0] << 8) | s[1]) == 0, "probe wants a stream without restart markers"
        elif m == 0xDA:
            for j in range(s[0]):
                h, v = frame[s[1 + 2 * j]]
                comps.append((h, v, s[2 + 2 * j] >> 4, s[2 + 2 * j] & 15))
            scan = bytes(data[i + seg:])
            end = scan.rfind(b"\xff\xd9")
            scan = scan[:end].replace(b"\xff\x00", b"\xff")
            return dc, ac, comps, np.unpackbits(np.frombuffer(scan, np.uint8))
        i += seg
    raise ValueError("no scan")


def run(bits, dc, ac, comps, start, true_states, max_symbols):
    """decode from bit `start` in state (block 0 of the MCU, k = 0).  true_states: None -> record and return the set of
    (bit position, block-in-MCU, k) at every symbol start; otherwise -> (bits, symbols) until the state is in it."""
    blocks = [c for c in comps for _ in range(c[0] * c[1])]  # the MCU's block sequence
    pos, b, k, n, seen = start, 0, 0, 0, set()
    nbits = len(bits)
    while pos < nbits and n < max_symbols:
        if true_states is None:
            seen.add((pos, b, k))
        elif (pos, b, k) in true_states:
            return pos - start, n
        table = (dc if k == 0 else ac)[blocks[b][2 if k == 0 else 3]]
        code, ln, sym = 0, 0, None
        while ln < 16 and pos < nbits:
            code = (code << 1) | int(bits[pos])
            pos += 1
            ln += 1
            sym = table.get((ln, code))
            if sym is not None:
                break
        if sym is None:  # not a code: a real decoder would carry on; count it as one symbol of 16 bits
            sym = 0
        n += 1
        if k == 0:
            pos += sym & 15
            k = 1
        else:
            r, s = sym >> 4, sym & 15
            pos += s
            k = 64 if (s == 0 and r != 15) else k + (16 if s == 0 else r + 1)
        if k >= 64:
            k = 0
            b = (b + 1) % len(blocks)
    return seen if true_states is None else None


def advance(bits, dc, ac, comps, state, stop_at):
    """decode from `state` = (bit position, block of the MCU, k) to the first symbol start at or after `stop_at`"""
    blocks = [c for c in comps for _ in range(c[0] * c[1])]
    pos, b, k = state
    nbits = len(bits)
    while pos < stop_at and pos < nbits:
        table = (dc if k == 0 else ac)[blocks[b][2 if k == 0 else 3]]
        code, ln, sym = 0, 0, None
        while ln < 16 and pos < nbits:
            code = (code << 1) | int(bits[pos])
            pos += 1
            ln += 1
            sym = table.get((ln, code))
            if sym is not None:
                break
        if sym is None:
            sym = 0
        if k == 0:
            pos += sym & 15
            k = 1
        else:
            r, s = sym >> 4, sym & 15
            pos += s
            k = 64 if (s == 0 and r != 15) else k + (16 if s == 0 else r + 1)
        if k >= 64:
            k = 0
            b = (b + 1) % len(blocks)
    return (pos, b, k)


def parallel_rounds(bits, dc, ac, comps, S):
    """the parallel scheme itself, run sequentially: sub-sequences of S bits, one decoder each.  Decoder i holds a
    candidate state for boundary i (the true one for i = 0, a guess -- block 0, k = 0 at bit i * S -- otherwise) and
    computes the state it implies at boundary i + 1; a round replaces every candidate by its left neighbour's
    implication.  Returns (rounds until nothing changes, whether the fixed point is the sequential decode)."""
    n = (len(bits) + S - 1) // S
    cand = [(i * S, 0, 0) for i in range(n)]
    truth = [(0, 0, 0)]
    for i in range(1, n):
        truth.append(advance(bits, dc, ac, comps, truth[-1], i * S))
    rounds = 0
    while True:
        implied = [advance(bits, dc, ac, comps, cand[i], (i + 1) * S) for i in range(n - 1)]
        new = [cand[0]] + implied
        rounds += 1
        if new == cand:
            break
        cand = new
    return rounds, cand == truth, n


def main():
    import cv2
    from tools import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--quality", type=int, default=75)
    ap.add_argument("--starts", type=int, default=400)
    ap.add_argument("--size", type=int, nargs=2, default=(640, 512))
    ap.add_argument("--subsequence", type=int, nargs="*", default=[1024, 4096],
                    help="also run the parallel scheme with sub-sequences of this many bits")
    a = ap.parse_args()
    seq = synth.StereoSequence(a.size[0], a.size[1], n_frames=1, tex_size=1024)
    ok, enc = cv2.imencode(".jpg", seq.frames[0][0], [cv2.IMWRITE_JPEG_QUALITY, a.quality])
    dc, ac, comps, bits = parse(enc.tobytes())
    truth = run(bits, dc, ac, comps, 0, None, 1 << 60)
    rs = np.random.RandomState(0)
    dist, syms, lost = [], [], 0
    for s in rs.randint(0, len(bits) - 20000, a.starts):
        r = run(bits, dc, ac, comps, int(s), truth, 3000)
        if r is None:
            lost += 1
        else:
            dist.append(r[0])
            syms.append(r[1])
    d, n = np.array(dist), np.array(syms)
    pct = lambda v, q: float(np.percentile(v, q)) if len(v) else None
    par = {}
    for S in a.subsequence:
        rounds, ok, ndec = parallel_rounds(bits, dc, ac, comps, S)
        par[str(S)] = {"decoders": ndec, "rounds_to_fixed_point": rounds, "equals_sequential_decode": ok}
    print(json.dumps({"image": f"{a.size[0]}x{a.size[1]} synthetic frame, q{a.quality}, 4:2:0 interleaved",
                      "scan_bits": int(len(bits)), "starts": a.starts, "not_locked_within_3000_symbols": lost,
                      "lock_in_bits": {"median": pct(d, 50), "p90": pct(d, 90), "p99": pct(d, 99), "max": pct(d, 100)},
                      "parallel_scheme": par,
                      "lock_in_symbols": {"median": pct(n, 50), "p90": pct(n, 90), "p99": pct(n, 99), "max": pct(n, 100)}}))


if __name__ == "__main__":
    main()
