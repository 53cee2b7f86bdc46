"""The CPU oracle against (a) the fixtures minted from the reference itself and (b) its own
internal consistency.  No GPU needed."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import cpu_oracle as O


def _cases(golden_dir):
    with open(os.path.join(golden_dir, "index_kat.json")) as f:
        return json.load(f)


def test_survey_kat():
    # SURVEY.md A.1 [probe]: grid=[[1,0,1,0],[0,1,1,0]]
    grid = torch.tensor([[1, 0, 1, 0], [0, 1, 1, 0]], dtype=torch.bool).view(1, 1, 2, 4)
    gi, me = O.grid_mappings(grid)
    assert gi.flatten().tolist() == [0, -8, 1, -7, -6, 2, 3, -5]
    assert me.tolist() == [0, 2, 5, 6]


def test_index_tensors_match_reference(golden_dir):
    """grid_idx / mapping_exec / transfer_idx == reference get_grid_mappings + _process_grid."""
    n = 0
    for case in _cases(golden_dir):
        prev_gi = None
        for fr in case["frames"]:
            grid = torch.tensor(fr["grid"], dtype=torch.bool).view(fr["shape"])
            gi, me = O.grid_mappings(grid)
            assert gi.flatten().tolist() == fr["grid_idx"], case["name"]
            assert me.tolist() == fr["mapping_exec"], case["name"]
            if fr["transfer_idx"] is not None:
                assert O.transfer_idx(grid, prev_gi).tolist() == fr["transfer_idx"], case["name"]
            prev_gi = gi
            n += 1
    assert n >= 20


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_split_combine_roundtrip(dtype):
    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 5, 12, 20, generator=g).to(dtype)
    grid = torch.rand(2, 1, 3, 5, generator=g) < 0.5
    gi, me = O.grid_mappings(grid)
    tiles = O.split(img, me, 4)
    # every tile is the slice of its cell
    for b, cell in enumerate(me.tolist()):
        n, r = divmod(cell, 15)
        gh, gw = divmod(r, 5)
        assert torch.equal(tiles[b], img[n, :, gh * 4:gh * 4 + 4, gw * 4:gw * 4 + 4])
    out = torch.zeros_like(img)
    O.combine_(tiles, out, me)
    mask = grid.repeat_interleave(4, 2).repeat_interleave(4, 3).expand_as(img)
    assert torch.equal(out[mask], img[mask]) and float(out[~mask].abs().sum()) == 0.0


@pytest.mark.parametrize("N,GH,GW,BS,C,pad", [(1, 4, 8, 8, 3, 1), (2, 3, 5, 4, 6, 1), (1, 2, 4, 8, 2, 3),
                                              (1, 3, 3, 4, 4, 2), (2, 2, 2, 2, 3, 1)])
def test_ring_protocol_equals_plane(N, GH, GW, BS, C, pad):
    """The equivalence the B200 design stands on: the reference's
    transfer(ring)+repad over tile FIFOs == cropping a persistent dense plane in which executed
    cells are overwritten in place.  Bit-exact over a seeded clip with poisoned interiors."""
    g = torch.Generator().manual_seed(GH * 100 + GW * 10 + pad)
    ring = O.RingProtocol()
    plane = torch.zeros(N, C, GH * BS, GW * BS, dtype=torch.float16)
    fracs = [1.0, 0.3, 0.5, 0.0, 0.2, 1.0, 0.6, 0.4]
    for t, frac in enumerate(fracs):
        grid = torch.rand(N, 1, GH, GW, generator=g) < frac if t else torch.ones(N, 1, GH, GW, dtype=torch.bool)
        gi, me = O.grid_mappings(grid)
        tiles = torch.randn(me.numel(), C, BS, BS, generator=g).to(torch.float16)
        expected = ring.step(tiles, grid, pad)
        O.combine_(tiles, plane, me)
        got = O.plane_halo(plane, me, BS, pad)
        assert not torch.isnan(expected.float()).any(), "a poisoned interior was consumed"
        assert torch.equal(got.view(torch.int16), expected.view(torch.int16)), f"frame {t}"


def test_oracle_matches_reference_cuda_kernels(golden_dir):
    """Pin: outputs of the reference's own CUDA C (NVRTC, sm_100a, B200) for the four kernels."""
    files = sorted(glob.glob(os.path.join(golden_dir, "ref_kernels_*.npz")))
    if not files:
        pytest.skip("ref_kernels_*.npz not generated yet (oracle/make_golden_gpu.py on the GPU box)")
    checked = 0
    for f in files:
        z = np.load(f)
        dt = torch.float16 if str(z["dtype"]) == "float16" else torch.float32
        T = lambda k: torch.from_numpy(z[k].copy())  # noqa: E731
        image, grid = T("image"), T("grid").bool()
        BS, pad = int(z["BS"]), int(z["pad"])
        gi, me = O.grid_mappings(grid)
        assert torch.equal(gi, T("grid_idx")) and torch.equal(me, T("mapping_exec"))
        tiles = O.split(image, me, BS)
        assert torch.equal(tiles.view(torch.int16 if dt == torch.float16 else torch.int32),
                           T("split").view(torch.int16 if dt == torch.float16 else torch.int32))
        out = T("combine_base").clone()
        O.combine_(T("tiles_in"), out, me)
        assert torch.equal(out, T("combine")), f
        tr = T("transfer_base").clone()
        O.transfer(tr, T("prev_exec"), T("prev_transfer"), T("transfer_idx"), grid.numel(), pad)
        assert torch.equal(tr, T("transfer")), f
        rp = O.repad(T("tiles_in"), T("transfer"), gi, me, pad)
        assert torch.equal(rp, T("repad")), f
        checked += 1
    assert checked == len(files)
