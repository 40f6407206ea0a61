"""z-slab building blocks on ONE GPU: the slabs of a decomposed mesh live side by side on the device, so the kernels that
read their halo planes from a neighbour's memory (`dzb_tm_dc_nodal_range_peer`, `dzb_tm_maxwell_range_peer(_sync)`) can
be checked against the undecomposed operator without a second GPU -- the neighbour's array is just another allocation.
(The multi-process versions, NCCL / symmetric memory, are `tools/check_slab.py` and `tests/test_slab_cpu.py`.)

In-kernel ordering (the sync words of dzb.h) cannot WAIT here -- two launches that wait for each other's start must not
share a device -- so its bookkeeping is checked with the neighbours' words pre-armed."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device (no CPU fallback exists)")


def _neighbours(parts, r):
    """(lower, upper) peer tuples of slab r when all slabs live on this device."""
    L = parts[r].layout
    field = lambda p: p.phi if hasattr(p, "phi") else p.e
    lower = upper = None
    if L.has_lower:
        Ll = parts[r - 1].layout
        lower = (field(parts[r - 1]).data_ptr(), Ll.nz_local, L.p_own0 - 1, Ll.p_own1 - 1)
    if L.has_upper:
        Lu = parts[r + 1].layout
        upper = (field(parts[r + 1]).data_ptr(), Lu.nz_local, L.p_own1, Lu.p_own0)
    return lower, upper


def _sync_words(world):
    """Per-slab sync words with the neighbours' announcements pre-armed (any epoch has been announced already)."""
    words = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(world)]
    for w in words:
        w[0] = 1 << 30
        w[1] = 1 << 30
    return words


@pytest.mark.parametrize("shape,world", [((40, 33, 29), 2), ((40, 33, 29), 3), ((41, 32, 30), 2), ((41, 32, 30), 3),
                                         ((70, 20, 19), 2), ((70, 20, 19), 3), ((37, 18, 5), 4), ((600, 9, 7), 7)])
def test_dc_slabs_with_peer_read_halos(shape, world):
    from discretize_b200.slab import SlabDCNodal

    nx, ny, nz = shape
    rng = np.random.default_rng(1)
    h = (rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(nz) + 0.5)
    whole = SlabDCNodal(nx, ny, nz, 0, 1, h=h)
    whole.step()
    parts = [SlabDCNodal(nx, ny, nz, r, world, h=h) for r in range(world)]
    words = _sync_words(world)
    for use_sync in (False, True):
        for launch in range(2):
            for r, p in enumerate(parts):
                L = p.layout
                lower, upper = _neighbours(parts, r)
                sync = None
                if use_sync:
                    lo_sig = words[r - 1][1:].data_ptr() if L.has_lower else 0
                    hi_sig = words[r + 1][0:].data_ptr() if L.has_upper else 0
                    sync = (words[r].data_ptr(), lo_sig, hi_sig)
                p.out.fill_(float("nan"))
                p.op.apply_planes_peer(p.phi, p.out, L.p_own0, L.p_own1, lower, upper, sync=sync)
                ref = whole.out[L.owned_node_slice()]
                assert torch.equal(p.owned_output(), ref), (shape, world, r, use_sync)
            if use_sync:
                torch.cuda.synchronize()
                for r, w in enumerate(words):
                    w = w.cpu().numpy()
                    assert w[2] == launch + 1 and w[3] == 0 and w[4] == 0          # epochs closed, nothing timed out
                    # what the neighbours' launches wrote into this slab's words: their epoch (overwrites the pre-arming)
                    if parts[r].layout.has_lower:
                        assert w[0] == launch + 1
                    if parts[r].layout.has_upper:
                        assert w[1] == launch + 1
                for w in words:            # re-arm for the next round of launches
                    w[0] = 1 << 30
                    w[1] = 1 << 30


@pytest.mark.parametrize("shape,world", [((40, 32, 29), 2), ((40, 32, 29), 3), ((64, 20, 18), 2), ((64, 20, 18), 3),
                                         ((62, 34, 12), 2), ((62, 34, 12), 3), ((36, 18, 5), 4), ((130, 10, 7), 7)])
def test_maxwell_slabs_with_peer_read_halos(shape, world):
    from discretize_b200.slab import SlabMaxwell

    nx, ny, nz = shape
    rng = np.random.default_rng(2)
    h = (rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(nz) + 0.5)
    whole = SlabMaxwell(nx, ny, nz, 0, 1, h=h)
    whole.step()
    esw = whole.es
    parts = [SlabMaxwell(nx, ny, nz, r, world, h=h) for r in range(world)]
    words = _sync_words(world)
    for use_sync in (False, True):
        for r, p in enumerate(parts):
            L = p.layout
            lower, upper = _neighbours(parts, r)
            sync = None
            if use_sync:
                lo_sig = words[r - 1][1:].data_ptr() if L.has_lower else 0
                hi_sig = words[r + 1][0:].data_ptr() if L.has_upper else 0
                sync = (words[r].data_ptr(), lo_sig, hi_sig)
            p.out.fill_(float("nan"))
            p.op.apply_planes_peer(p.e, p.out, L.p_own0, L.p_own1, lower, upper, sync=sync)
            g0, g1 = L.c0, L.c1 + (0 if L.has_upper else 1)
            ref = (whole.out[g0 * esw.px:g1 * esw.px], whole.out[esw.off_y + g0 * esw.py:esw.off_y + g1 * esw.py],
                   whole.out[esw.off_z + L.c0 * esw.pz:esw.off_z + L.c1 * esw.pz])
            for got, want in zip(p.owned_parts(p.out), ref):
                assert float((got - want).abs().max()) <= 1e-14 * float(want.abs().max()), (shape, world, r, use_sync)
        if use_sync:
            torch.cuda.synchronize()
            for w in words:
                w = w.cpu().numpy()
                assert w[2] == 1 and w[3] == 0 and w[4] == 0
