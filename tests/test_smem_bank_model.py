"""Host model of the shared-memory bank behaviour of the per-line Stockham stages (fsm_fft.cuh) for the three-stage
decompositions: 64-bit accesses are served per half-warp, conflict-free when the 16 lanes touch 16 distinct bank pairs
(slot index mod 16). Restates the index functions of line_fft_head / fft_last_item; documents why the last stage's input
has its own layout (pad_last) when R0 = R1 = 8 (profiles/r2f_phys3d_shared_wavefronts.txt)."""
import pytest


def pad(i, shift=4):
    return i + (i >> shift)


def pad_last(i, xtra):
    return pad(i) + xtra * (((i >> 6) + 1) >> 1)


def half_warps(tl):
    lanes = list(range(min(tl, 64)))
    return [lanes[i:i + 16] for i in range(0, len(lanes), 16)]


def conflict_free(slots):
    return len({s % 16 for s in slots}) == len(slots)


@pytest.mark.parametrize("n,ept,r0,r1,r2", [(512, 8, 8, 8, 8), (256, 8, 8, 8, 4), (1024, 16, 16, 8, 8), (512, 16, 16, 8, 4)])
def test_three_stage_exchanges_are_conflict_free(n, ept, r0, r1, r2):
    tl = n // ept
    xtra = 4 if (r0 == 8 and r1 == 8) else 0
    ns = r0
    for hw in half_warps(tl):
        if len(hw) < 16:
            continue
        for q in range(ept // r0):                      # stage 0 stores: buf[pad(w * R0 + t)], w = tau + q TL
            for t in range(r0):
                assert conflict_free([pad((tau + q * tl) * r0 + t) for tau in hw]), "stage 0 store"
        for q in range(ept // r1):
            for t in range(r1):                         # stage 1 loads: buf[pad(w + t N/R1)]
                assert conflict_free([pad(tau + q * tl + t * (n // r1)) for tau in hw]), "stage 1 load"
            for t in range(r1):                         # stage 1 stores: buf[pad_last((w / Ns) Ns R1 + (w % Ns) + t Ns)]
                slots = []
                for tau in hw:
                    w = tau + q * tl
                    slots.append(pad_last((w // ns) * ns * r1 + (w % ns) + t * ns, xtra))
                assert conflict_free(slots), "stage 1 store"
        rl, nsl = r2, n // r2
        for q in range(ept // rl):                      # last stage loads: buf[pad_last(w + t N/RL)]
            for t in range(rl):
                assert conflict_free([pad_last(tau + q * tl + t * nsl, xtra) for tau in hw]), "last stage load"


def test_plain_padding_conflicts_in_the_middle_stage_of_8x8():
    """The measured 2-way conflict: without the extra slots the two 8-lane groups of a half-warp land 68 slots apart."""
    n, r1, ns = 512, 8, 8
    hw = list(range(16))
    slots = [pad((tau // ns) * ns * r1 + (tau % ns)) for tau in hw]
    assert not conflict_free(slots)
    assert len({s % 16 for s in slots}) == 12           # 4 of 16 bank pairs collide


def test_pad_last_stays_inside_the_line_buffer_and_is_injective():
    for n, xtra in ((512, 4), (256, 4)):
        rawlen = n + (n >> 4) + 1 + xtra * (((n >> 6) + 1) >> 1)
        idx = [pad_last(i, xtra) for i in range(n)]
        assert len(set(idx)) == n and max(idx) < rawlen
