"""BASELINE config 5 ("tests/Cookie full pipeline: GPU DivQuant feeding host SRM + superpixel merge, segmentation equality
vs reference") as far as it can be built here: the reference's segmentation host code needs OpenCV's C++ libraries and
Xcode, so the harness is Python + the compiled reference pieces, stage for stage as
ClusteringSegmentationMain.cpp:124-375 runs them on tests/Cookie:

  srmMultiSegment -> generateSRM(Q = 128) -> SRM()      ClusteringSegmentation.cpp:8819-8836, 225-266
      GPU: the C4 edge list + bucket sort (dq_srm_sorted_edges) -> the REFERENCE's union-find merge loop, small-region merge
      and finalize (oracle/_ref/libsrm_ref.so: srm.c:179-190, 275-320) -> segmentation == the all-reference SRM()
  genHistogramsForBlocks                                 ClusteringSegmentation.cpp:365-569
      GPU: Vec3BToUID packing -> map_colors_mps on the 125-colour grid -> 4x4 block majority vote (dq_quant_blocks)
      == the two images the reference dumps (block_quant_full_output / block_quant_output), computed by the compiled
      reference's map_colors_mps and the oracle's unordered_map vote

The tags the live binary writes do not depend on DivQuant output at all (captureRegion returns before its
DivQuant calls), so these two products plus the SRM segmentation are everything of the pipeline the GPU path can change."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _bgr(px2d):
    return np.stack([px2d & 0xFF, (px2d >> 8) & 0xFF, (px2d >> 16) & 0xFF], axis=-1).astype(np.uint8)


@pytest.mark.parametrize("name", ["cookie", "batman"])
def test_pipeline_front_half_equals_all_reference_run(dq, oracle, reference, golden, name):
    import os
    from oracle import ReferenceSRM
    z = np.load(os.path.join(GOLDEN, f"{name}_px.npz"))
    h, w = int(z["shape"][0]), int(z["shape"][1])
    px2d = z["px"].reshape(h, w)          # 0x00RRGGBB = Vec3BToUID (superpixels/OpenCVUtil.h:18-27)
    image = _bgr(px2d)                    # the interleaved B,G,R bytes generateSRM hands to SRM()
    rs = ReferenceSRM()
    # ---- SRM: all-reference run vs GPU edges + reference merge ----
    ref_pairs, ref_seg = rs.sorted_edges(image, q=128.0)
    gpu_pairs = dq.srm_sorted_edges(image)
    assert np.array_equal(gpu_pairs, ref_pairs)
    seg = rs.run_with_pairs(image, gpu_pairs, q=128.0)
    assert np.array_equal(seg, ref_seg)
    assert len(np.unique(seg.reshape(-1, 3), axis=0)) > 1      # a real segmentation, not a constant image
    # ---- genHistogramsForBlocks: quantized image and block image ----
    grid = golden["grid125"]
    quant, blocks = dq.quant_blocks(px2d.ravel(), w, h, grid, 4)
    ref_quant = reference.map_colors_mps(px2d.ravel(), grid)
    assert np.array_equal(quant, ref_quant)
    assert np.array_equal(blocks, oracle.block_vote(ref_quant, w, h, 4))
    # and the label image the dead-code consumers ask for (mapQuantPixelsToColortableIndexes, OpenCVUtil.cpp:787-849)
    labels = dq.colortable_indexes(quant, grid)
    assert np.array_equal(grid[labels] & 0xFFFFFF, quant)
