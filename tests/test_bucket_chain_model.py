"""The estimator's bucket chains (rad_team_b200/csrc/rt_env_kernels.cuh: bucket_insert_median), as an executable
model on the CPU.

The CUDA kernel keeps a cell's readings in 32-byte buckets — seven ascending values + a link word (next bucket in
the low half; in the first bucket, the number of readings in the high half) — and inserts a reading by bubbling
it into its bucket and cascading the evicted maximum down the chain, picking the median on the way.  This file
restates that procedure step for step in Python (same loop, same early exit, same slot allocation by time index)
and checks it against `statistics.median` — the reference's `IntensityResamplingEstimator.get_estimate`
(src/math_utils/estimator.py:61-70) — and the chain invariants, over long random sequences with many ties.  The
kernel itself is pinned on the GPU against the oracle (tests/test_gpu_env_parity.py, test_gpu_randomized.py)."""
import random
import statistics

CAP = 7


class Pool:
    def __init__(self):
        self.b = {}        # bucket id -> [v0..v6, link]
        self.head = {}     # cell -> bucket id + 1

    def insert(self, cell, idx, cnt):
        cur = self.head.get(cell, 0)
        if cur == 0:
            self.b[idx] = [cnt, 0, 0, 0, 0, 0, 0, 1 << 16]
            self.head[cell] = idx + 1
            return float(cnt), 1, 0
        bp = cur - 1
        n = (self.b[bp][7] >> 16) + 1
        k1, k2 = (n - 1) >> 1, n >> 1
        m1 = m2 = 0
        carry, pending, base, left = cnt, True, 0, n - 1
        hops = 0
        for hop in range((n + 5) // 7 + 1):
            hops += 1
            v = list(self.b[bp][:7])
            nxt = self.b[bp][7] & 0xFFFF
            count_word = (n << 16) if hop == 0 else 0
            m = min(left, CAP)
            left -= m
            m_new, dirty = m, hop == 0
            if pending and (m < CAP or carry < v[6]):
                x = carry
                for i in range(CAP):
                    if i < m and v[i] > x:
                        v[i], x = x, v[i]
                if m < CAP:
                    v[m] = x
                    m_new, pending = m + 1, False
                else:
                    carry = x
                dirty = True
            for i in range(m_new):
                if base + i == k1:
                    m1 = v[i]
                if base + i == k2:
                    m2 = v[i]
            base += m_new
            if nxt == 0 and pending:
                self.b[idx] = [carry, 0, 0, 0, 0, 0, 0, 0]
                if base == k1:
                    m1 = carry
                if base == k2:
                    m2 = carry
                nxt, pending, dirty, left = idx + 1, False, True, -1
            if dirty:
                self.b[bp] = v + [nxt | count_word]
            if left < 0 or nxt == 0 or (not pending and base > k2):
                break
            bp = nxt - 1
        med = float(m2) if n & 1 else (m1 + m2) / 2.0
        return med, n, hops

    def chain(self, cell):
        out, bp, n = [], self.head[cell] - 1, None
        while True:
            row = self.b[bp]
            if n is None:
                n = row[7] >> 16
            take = min(CAP, n - len(out))
            out += row[:take]
            if (row[7] & 0xFFFF) == 0:
                break
            bp = (row[7] & 0xFFFF) - 1
        return out, n


def run(seed, n_cells, steps, lam_like):
    rng = random.Random(seed)
    pool, truth = Pool(), {}
    for t in range(steps):
        cell = rng.randrange(n_cells)
        cnt = max(0, int(rng.gauss(lam_like, lam_like ** 0.5)))      # clustered values, many ties
        truth.setdefault(cell, []).append(cnt)
        med, n, hops = pool.insert(cell, t, cnt)
        assert n == len(truth[cell])
        assert med == statistics.median(truth[cell]), (seed, t, cell)
        assert hops <= (n + CAP - 1) // CAP                            # ceil(n / 7) dependent loads, never more
    for cell, vals in truth.items():
        got, n = pool.chain(cell)
        assert n == len(vals) and got == sorted(vals)                  # whole chain ascending, nothing lost


def test_bucket_chain_median_equals_statistics_median():
    run(1, n_cells=1, steps=400, lam_like=30.0)        # one cell: a chain of 58 buckets, cascades on every insert
    run(2, n_cells=4, steps=600, lam_like=12.0)        # 2 x 2 grid, heavy revisits
    run(3, n_cells=64, steps=720, lam_like=200.0)      # typical episode: most cells stay within one bucket
    run(4, n_cells=3, steps=300, lam_like=0.6)         # almost all zeros and ones


def test_bucket_ids_fit_the_cell_record():
    # a bucket is opened at most once per (tick, agent): ids stay below T*A <= 65534, so id + 1 fits 16 bits
    pool = Pool()
    for t in range(65534):
        pool.insert(t % 2053, t, t & 7)
    assert max(pool.b) + 1 <= 0xFFFF
    for cell in range(2053):
        got, n = pool.chain(cell)
        assert got == sorted(got) and n == len(got)
