"""GPU: gs_bin_prepare -- the hand-written depth sort (csrc/depthsort.cu; reference _sort_gaussians_by_depth,
src/core/renderer.py:222-239) -- alone, through the C ABI, against torch's stable sort on the same keys: every size class
(one item, partial tile, exactly one tile, tile + 1, several tiles per CTA), key ranges that need 1, 2, 3 and 4 eight-bit
passes, heavy ties, no valid key at all.  Bit-exact order, prefix sums and counters."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _prepare(keys_u32: torch.Tensor, tt: torch.Tensor):
    from importlib import import_module
    _lib = import_module("mini-3d-gaussian-splatting_b200._lib")
    lib = _lib.load()
    n = keys_u32.numel()
    P = _lib.ptr
    counters = torch.full((3,), -1, dtype=torch.int64, device="cuda")
    sorted_ids = torch.full((max(n, 1),), -1, dtype=torch.int32, device="cuda")
    offsets = torch.full((max(n, 1),), -1, dtype=torch.int64, device="cuda")
    wsb = int(lib.gs_bin_workspace_bytes(n, 0, 64))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.gs_bin_prepare(n, P(keys_u32), P(tt), P(ws), wsb, P(sorted_ids), P(offsets), P(counters), st), "gs_bin_prepare")
    torch.cuda.synchronize()
    return sorted_ids[:n], offsets[:n], counters


def _check(keys_i64: torch.Tensor, tt: torch.Tensor, what):
    """keys_i64: the unsigned 32-bit keys held in int64 (torch has no uint32 sort)."""
    n = keys_i64.numel()
    keys_u32 = (keys_i64 & 0xFFFFFFFF).to(torch.int64)
    dev_keys = torch.where(keys_u32 >= 2 ** 31, keys_u32 - 2 ** 32, keys_u32).to(torch.int32).cuda()     # same bit patterns
    tt = torch.where(keys_u32 >= 0xFFFFFFFE, torch.zeros_like(tt), tt)          # splats without tiles touch none
    sorted_ids, offsets, counters = _prepare(dev_keys, tt.to(torch.int32).cuda())
    order = torch.sort(keys_u32, stable=True).indices                           # ties -> ascending index
    assert torch.equal(sorted_ids.cpu().long(), order), what
    w = tt.long()[order]
    want_off = torch.cumsum(w, 0) - w
    assert torch.equal(offsets.cpu(), want_off), what
    assert counters.tolist() == [int((keys_u32 < 0xFFFFFFFE).sum()), int(tt.sum()), int((keys_u32 < 0xFFFFFFFF).sum())], what


@pytest.mark.parametrize("n", [1, 2, 31, 33, 80, 1791, 1792, 1793, 7167, 7168, 7169, 50_000, 265_217, 1_000_000, 3_000_001])
def test_depth_order_matches_stable_sort_for_every_size_class(n):
    g = torch.Generator().manual_seed(n)
    z = 1.27 + 3.5 * torch.rand(n, generator=g)                   # the depth range of the bench scene: 3 passes
    keys = z.view(torch.int32).long()
    r = torch.rand(n, generator=g)
    keys[r < 0.05] = 0xFFFFFFFF                                   # culled
    keys[(r >= 0.05) & (r < 0.06)] = 0xFFFFFFFE                   # visible, empty AABB
    tt = torch.randint(1, 65, (n,), generator=g)
    _check(keys, tt, f"n={n}")


@pytest.mark.parametrize("case", ["one_binade_8bit", "16bit", "24bit", "31bit", "all_equal", "heavy_ties", "no_valid_key", "only_valid",
                                  "huge_rect_counts"])
def test_depth_order_key_ranges_and_degenerate_inputs(case):
    n = 123_457
    g = torch.Generator().manual_seed(7)
    tt = torch.randint(1, 65, (n,), generator=g)
    base = torch.tensor(2.0).view(torch.int32).long()
    if case == "one_binade_8bit":
        keys = base + torch.randint(0, 200, (n,), generator=g)
    elif case == "16bit":
        keys = base + torch.randint(0, 60000, (n,), generator=g)
    elif case == "24bit":
        keys = base + torch.randint(0, 1 << 24, (n,), generator=g)
    elif case == "31bit":                                         # depths from 1e-30 to 1e30: four passes
        keys = (10.0 ** (60 * torch.rand(n, generator=g) - 30)).to(torch.float32).view(torch.int32).long()
    elif case == "all_equal":
        keys = base.repeat(n)
    elif case == "heavy_ties":
        keys = base + torch.randint(0, 7, (n,), generator=g) * 1000
    elif case == "no_valid_key":
        keys = torch.where(torch.rand(n, generator=g) < 0.5, torch.tensor(0xFFFFFFFF), torch.tensor(0xFFFFFFFE)).long()
    elif case == "only_valid":
        keys = (0.5 + torch.rand(n, generator=g)).view(torch.int32).long()
    else:
        keys = (1.0 + torch.rand(n, generator=g)).view(torch.int32).long()
        tt = torch.randint(1, 8161, (n,), generator=g)            # a splat may cover the whole 120x68 grid
    if case not in ("no_valid_key", "only_valid"):
        r = torch.rand(n, generator=g)
        keys = torch.where(r < 0.1, torch.tensor(0xFFFFFFFF), keys)
        keys = torch.where((r >= 0.1) & (r < 0.12), torch.tensor(0xFFFFFFFE), keys)
    _check(keys, tt, case)


def test_depth_order_is_deterministic_and_reentrant():
    n = 400_000
    g = torch.Generator().manual_seed(3)
    keys = (1.0 + 4 * torch.rand(n, generator=g)).view(torch.int32).cuda()
    tt = torch.randint(0, 30, (n,), generator=g).to(torch.int32).cuda()
    a = _prepare(keys, tt)
    for _ in range(3):
        b = _prepare(keys, tt)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
