"""CPU self-consistency of the sub-box parity checker (oracle/subbox_check.py): the
oracle's own full result of a chunked stack, fed in as if it were the GPU's, must
compare clean against the oracle recomputation of every core - i.e. the margins the
checker leaves around a core (block grid, filter radius, two pruning hops) suffice."""
import numpy as np

from oracle import magmap_restated as mm
from oracle import subbox_check as sb
from magellanmapper_b200 import synth


def test_cores_of_a_chunked_stack_reproduce_the_full_oracle():
    shape = (60, 130, 120)
    vol, _ = synth.make_volume(shape, seed=11, density=1 / 2500.0)
    nm = synth.near_max_of(vol)
    prof = mm.Profile(segment_size=50)
    final, seg_rois, blocks = mm.detect_blobs_blocks(vol, prof, (1, 1, 1), nm, return_parts=True)
    assert blocks.sub_roi_slices.shape == (2, 3, 3)

    def fetch(z0, z1, y0, y1, x0, x1):
        return vol[z0:z1, y0:y1, x0:x1]

    out = sb.check_stack(fetch, shape, prof, (1, 1, 1), nm, seg_rois, final, processes=2,
                         core_size=24)
    assert out["boxes"] == 5 and out["oracle_blobs"] > 0
    assert out["unexplained_diff"] == 0 and out["f1"] == 1.0, out
    assert out["seam_rows_differing"] == 0 and out["seam_rows_gpu"] == len(final)
    # a corrupted table is noticed: drop one blob of the first core's chunk
    tab = seg_rois[0, 0, 0]
    in_core = np.all(tab[:, :3] < 24, axis=1)
    if in_core.any():
        seg_bad = seg_rois.copy()
        seg_bad[0, 0, 0] = tab[~(in_core & (np.cumsum(in_core) == 1))]
        bad = sb.check_stack(fetch, shape, prof, (1, 1, 1), nm, seg_bad, final, core_size=24)
        assert bad["unexplained_diff"] + bad["near_threshold"] + bad["order_dependent"] >= 1


def test_region_is_block_aligned_and_clipped_to_the_chunk():
    region, valid = sb.region_for_core(((100, 148), (0, 48), (457, 505)), (505, 505, 505),
                                       (25, 25, 25), 20, 5.0)
    assert region[0] == (25, 225) and region[1] == (0, 125) and region[2] == (375, 505)
    assert valid[1][0] == 0 and valid[2][1] == 505          # chunk faces stay exact
    assert valid[0] == (46, 204)
