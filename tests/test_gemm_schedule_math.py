"""CPU checks of the host / device integer arithmetic behind the GEMM tile schedule (csrc/gemm_tc.cu), transliterated
line by line: the multiply-high division every role decodes tile indices with, the two decode orders (row-tile major /
column-block major for the BN-statistics launches, CTA pairs sharing a schedule slot), the static round robin, and the
counter windows of the dynamic schedule (fetches per launch = chunks + one end marker per CTA).  The kernels themselves
need a GPU: tests/test_gpu_kernels.py."""
import random


def make_fastdiv(d):
    """make_fastdiv(): mul == 0 encodes d == 1."""
    if d <= 1:
        return 0, 0
    l = 0
    while (1 << l) < d:
        l += 1
    pw = 31 + l
    m = ((1 << pw) + d - 1) // d
    assert (1 << 31) <= m < (1 << 32)      # never 0 (0 encodes d == 1); exactly 2^31 for powers of two
    return m, pw - 32


def fdiv(x, f):
    mul, shr = f
    return x if mul == 0 else ((x * mul) >> 32) >> shr


def test_multiply_high_division_is_exact_over_the_schedule_range():
    rng = random.Random(0)
    divisors = list(range(1, 600)) + [rng.randrange(600, 1 << 20) for _ in range(400)] + [6272, 25088, 100352, (1 << 20) - 1]
    for d in divisors:
        f = make_fastdiv(d)
        xs = [0, 1, d - 1, d, d + 1, 2 * d - 1, 7 * d, (1 << 31) - 1, ((1 << 31) - 1) // d * d, ((1 << 31) - 1) // d * d - 1]
        xs += [rng.randrange(0, 1 << 31) for _ in range(50)]
        for x in xs:
            if 0 <= x < (1 << 31):
                assert fdiv(x, f) == x // d, (x, d)


def decode_tile(t, m_sched, n_tiles, nt_major, pair, rank):
    """decode_tile(): schedule index -> (split, row tile, column tile)."""
    d_mn, d_nt, d_mt = make_fastdiv(m_sched * n_tiles), make_fastdiv(n_tiles), make_fastdiv(m_sched)
    ks = fdiv(t, d_mn)
    rem = t - ks * (m_sched * n_tiles)
    if nt_major:
        nt = fdiv(rem, d_mt)
        mt = rem - nt * m_sched
    else:
        mt = fdiv(rem, d_nt)
        nt = rem - mt * n_tiles
    if pair:
        mt = 2 * mt + rank
    return ks, mt, nt


def test_every_tile_is_decoded_exactly_once_in_both_orders_and_in_pairs():
    for m_tiles, n_tiles, k_splits in ((157, 1, 1), (392, 4, 1), (98, 8, 2), (61, 3, 4), (1, 1, 1)):
        for nt_major in (0, 1):
            for pair in (0, 1):
                m_sched = (m_tiles + 1) // 2 if pair else m_tiles
                total = m_sched * n_tiles * k_splits
                seen = set()
                for t in range(total):
                    for rank in ((0, 1) if pair else (0,)):
                        ks, mt, nt = decode_tile(t, m_sched, n_tiles, nt_major, pair, rank)
                        assert 0 <= ks < k_splits and 0 <= nt < n_tiles
                        if mt < m_tiles:      # the second CTA of the last pair of an odd row-tile count is all padding
                            assert (ks, mt, nt) not in seen
                            seen.add((ks, mt, nt))
                assert len(seen) == m_tiles * n_tiles * k_splits
                if nt_major and k_splits == 1:   # a CTA walking increasing indices changes its column block <= n_tiles - 1 times
                    nts = [decode_tile(t, m_sched, n_tiles, 1, pair, 0)[2] for t in range(total)]
                    assert nts == sorted(nts)


def producer_sequence(first, fetch, total, chunk):
    """The producer warp's loop: tile indices it publishes until (and excluding) its first end marker."""
    out, t = [], first
    while t < total:
        out.append(t)
        t = t + 1 if ((t + 1) % chunk != 0 and t + 1 < total) else fetch()
    return out


def test_dynamic_schedule_windows_hand_out_every_tile_once_with_a_fixed_number_of_fetches():
    rng = random.Random(1)
    base = 0xFFFFFF00          # the counters only ever grow: windows wrap modulo 2^32
    for total, grid, chunk in ((157, 148, 1), (1954, 148, 2), (5469, 148, 4), (7, 7, 1), (25088, 148, 4), (300, 148, 1)):
        counter = [base]

        def fetch():
            v = counter[0]
            counter[0] = (counter[0] + 1) & 0xFFFFFFFF
            return ((v - base) & 0xFFFFFFFF) * chunk

        # CTAs interleave arbitrarily: simulate by letting a random CTA take its next step until all have seen an end marker
        state = [{"t": fetch(), "done": False, "tiles": []} for _ in range(grid)]
        while not all(s["done"] for s in state):
            s = rng.choice([s for s in state if not s["done"]])
            t = s["t"]
            if t >= total:
                s["done"] = True
                continue
            s["tiles"].append(t)
            s["t"] = t + 1 if ((t + 1) % chunk != 0 and t + 1 < total) else fetch()
        tiles = sorted(t for s in state for t in s["tiles"])
        assert tiles == list(range(total))
        fetches = (counter[0] - base) & 0xFFFFFFFF
        assert fetches == (total + chunk - 1) // chunk + grid      # what the host adds to the counter's window base
        base = counter[0]


def test_static_round_robin_covers_every_tile_once():
    for total, workers in ((157, 148), (74, 74), (1000, 74), (3, 148)):
        grid = min(total, workers)
        tiles = sorted(t for c in range(grid) for t in range(c, total, grid))
        assert tiles == list(range(total))
