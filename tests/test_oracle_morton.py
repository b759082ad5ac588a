"""The reference's own Morton known-answer tests (src/sph/morton.rs:189-251), restated against the oracle."""
import numpy as np

from oracle import pyoracle as po

BIG_X, BIG_Y, BIG_M = 0b1111_0001_0010_0000, 0b1001_1101_1000_1100, 0b1101_0111_1010_0011_1000_0100_1010_0000


def test_encode_lookup_works_for_examples():  # morton.rs:190-198
    L = po.lib()
    assert L.yo_morton_encode_lookup(2, 2) == 12
    assert L.yo_morton_encode_lookup(3, 6) == 45
    assert L.yo_morton_encode_lookup(4, 0) == 16
    assert L.yo_morton_encode_lookup(BIG_X, BIG_Y) == BIG_M


def test_encode_bitfiddle_works_for_examples():  # morton.rs:201-209
    L = po.lib()
    assert L.yo_morton_encode(2, 2) == 12
    assert L.yo_morton_encode(3, 6) == 45
    assert L.yo_morton_encode(4, 0) == 16
    assert L.yo_morton_encode(BIG_X, BIG_Y) == BIG_M


def test_decode_bitfiddle_works_for_examples():  # morton.rs:216-228
    L = po.lib()
    for m, x, y in [(12, 2, 2), (45, 3, 6), (16, 4, 0), (BIG_M, BIG_X, BIG_Y)]:
        assert L.yo_morton_decode_x(m) == x
        assert L.yo_morton_decode_y(m) == y


def test_bigmin_jumps_to_next_pos_in_rect():  # morton.rs:235-241
    L = po.lib()
    for cur in (16, 19, 29, 35):
        assert L.yo_find_bigmin(cur, 12, 45) == 36


def test_bigmin_within_rect_gives_next_in_rect():  # morton.rs:244-246
    assert po.lib().yo_find_bigmin(14, 12, 45) == 15


def test_bigmin_at_border_of_section_gives_next_in_rect():  # morton.rs:249-251
    assert po.lib().yo_find_bigmin(15, 12, 45) == 36


def test_lookup_equals_bitfiddle_and_roundtrip():
    L = po.lib()
    rng = np.random.default_rng(1)
    xs = rng.integers(0, 65536, 2000)
    ys = rng.integers(0, 65536, 2000)
    for x, y in zip(xs, ys):
        m = L.yo_morton_encode(int(x), int(y))
        assert m == L.yo_morton_encode_lookup(int(x), int(y))
        assert L.yo_morton_decode_x(m) == x and L.yo_morton_decode_y(m) == y


def test_bigmin_against_bruteforce():
    """BIGMIN(cur) = smallest Morton code > cur that lies inside the rectangle (Tropf & Herzog)."""
    L = po.lib()
    rng = np.random.default_rng(2)
    for _ in range(200):
        x0, y0 = int(rng.integers(0, 12)), int(rng.integers(0, 12))
        x1, y1 = x0 + int(rng.integers(0, 4)), y0 + int(rng.integers(0, 4))
        mn, mx = L.yo_morton_encode(x0, y0), L.yo_morton_encode(x1, y1)
        inside = sorted(L.yo_morton_encode(x, y) for x in range(x0, x1 + 1) for y in range(y0, y1 + 1))
        for cur in range(mn, mx):
            if L.yo_is_in_rect(cur, mn, mx):
                continue  # the search only calls it for codes outside the rectangle (neighborhood_search.rs:215-221)
            expect = next(m for m in inside if m > cur)
            assert L.yo_find_bigmin(cur, mn, mx) == expect, (cur, mn, mx)
