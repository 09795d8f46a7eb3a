"""Prototype of the event-list formulation of the filter (numpy/python), checked against the oracle.

Per output row: every presence chain of a sample value along the columns (gaps <= 2r+1 keep a chain alive) is one
"event" (value, birth column c_b, death column c_d = last presence column + 2r+1); events sorted by
(c_b, vertical chain start of the value in column c_b). Pixel x lists the events with c_b <= x+2r < c_d in that order.
"""
import sys

import numpy as np

sys.path.insert(0, '/root/repo')
import oracle  # noqa: E402


def events_filter(m, W, H, nn, r):
    span = 2 * r + 1
    sx, sy = W * (nn[0] // 2), H * (nn[1] // 2)
    reg = m[sy - r:sy + H + r, sx - r:sx + W + r].astype(np.int64)
    PH, PW = reg.shape
    vstart = np.zeros_like(reg)
    for c in range(PW):
        last = {}
        start = {}
        for p in range(PH):
            s = int(reg[p, c])
            if s not in last or p - last[s] > span:
                start[s] = p
            last[s] = p
            vstart[p, c] = start[s]
    items = []
    weights = []
    offs = [0]
    inv = np.float32(1.0) / np.float32(span * span)
    for y in range(H):
        band = reg[y:y + span, :]
        # presence per column
        pres = [set(band[:, c].tolist()) for c in range(PW)]
        ev = []  # [s, cb, cd, vkey]
        open_ev = {}
        last = {}
        for c in range(PW):
            for s in pres[c]:
                if s not in last or c - last[s] > span:
                    if s in open_ev:
                        ev[open_ev[s]][2] = last[s] + span
                    p = y + 2 * r
                    while reg[p, c] != s:
                        p -= 1
                    open_ev[s] = len(ev)
                    ev.append([s, c, None, int(vstart[p, c])])
                last[s] = c
        for s, i in open_ev.items():
            ev[i][2] = last[s] + span
        ev.sort(key=lambda e: (e[1], e[3]))
        # horizontal window counts by 2D prefix
        for x in range(W):
            c = x + 2 * r
            win = reg[y:y + span, x:x + span]
            for s, cb, cd, _ in ev:
                if cb <= c < cd:
                    cnt = int((win == s).sum())
                    assert cnt > 0, (x, y, s, cb, cd)
                    items.append(s)
                    weights.append(np.float32(cnt) * inv)
            offs.append(len(items))
    return np.array(items, np.uint16), np.array(weights, np.float32), np.array(offs, np.uint32)


if __name__ == '__main__':
    rng = np.random.default_rng(1)
    for t in range(60):
        W = int(rng.integers(4, 24))
        H = int(rng.integers(4, 20))
        r = int(rng.integers(1, min(W, H) // 2 + 1)) * 2
        r = min(r, (min(W, H) // 2) * 2)
        if r == 0:
            r = 2
        B = int(rng.integers(1, 30))
        kind = t % 4
        if kind == 0:
            m = rng.integers(0, B, (3 * H, 3 * W))
        elif kind == 1:
            bs = int(rng.integers(2, 8))
            m = rng.integers(0, B, (3 * H // bs + 1, 3 * W // bs + 1)).repeat(bs, 0).repeat(bs, 1)[:3 * H, :3 * W]
        elif kind == 2:
            m = np.where(rng.random((3 * H, 3 * W)) < 0.9, 0, rng.integers(0, B, (3 * H, 3 * W)))
        else:
            period = 2 * r + 3
            m = np.zeros((3 * H, 3 * W), np.int64)
            m[:, ::period] = rng.integers(1, B + 1, (3 * H, len(range(0, 3 * W, period))))
        m = m.astype(np.uint16)
        a = oracle.run_port(m, (W, H), (3, 3), r)
        b = events_filter(m, W, H, (3, 3), r)
        ok = all((x.shape == y.shape and (x.view(np.uint8) == y.view(np.uint8)).all()) for x, y in zip(a, b))
        print(t, W, H, r, B, kind, ok)
        assert ok
