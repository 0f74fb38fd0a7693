"""CPU oracle for the few-shot detection training hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of
the reference's hot path (bingykang/Fewshot_Detection @ /root/reference):
`darknet_meta.Darknet.forward`, `darknet.Darknet.forward`,
`dynamic_conv.DynamicConv2d`, `pooling.GlobalMaxPool2d`,
`region_loss.{neg_filter, build_targets, RegionLoss, RegionLossV2}` and
`utils.{bbox_iou, bbox_ious}`, and of the rows SURVEY.md 8f widens into:
`utils.{get_region_boxes, get_region_boxes_v2, nms}` + the reweight ensembling
and result-line format of valid_ensemble.py (oracle/utils.py), and
`image.data_augmentation` with Pillow's uint8 resize / HSV / point algorithms
restated in numpy (oracle/image.py).  Each function cites the reference
file:line it follows.

Who may import it: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` — as the checker or as the timed CPU
baseline, never as a product path.  Nothing under `fewshot_detection_b200/`
imports it; the product fails loudly when `libfsdet.so` is missing.

Where the arithmetic really lives: the reference delegates all tensor math to
the un-vendored third-party dependency torch==0.3.1 (requirements.txt:3).  The
oracle uses torch 2.11 CPU ops with the same documented semantics for the
network layers, and plain numpy / Python floats for `build_targets`, following
torch-0.3.1's behaviour that indexing a tensor down to one element yields a
Python float (so phase 2 of `build_targets` runs in float64).

Pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c).  The oracle is pinned instead against outputs of THE
REFERENCE ITSELF run in the build container: tests/golden/make_golden.py imports
the reference's own `region_loss.py`, `darknet_meta.py`, `darknet.py`,
`utils.py`, `cfg.py`, `dynamic_conv.py`, `pooling.py` from /root/reference
(made runnable under Py3/torch-2 by in-memory mechanical substitutions listed in
that script) and writes tests/golden/*.npz; tests/test_oracle_golden.py checks
every oracle function against those files (bit-exact for masks / indices /
counters / float32 IoUs, <=1e-6 relative for float tensors).  Likewise
make_golden_detect.py -> detect.npz (tests/test_oracle_detect.py, bit-exact)
and make_golden_augment.py -> augment.npz (the reference's image.py runs
unmodified under Pillow 12.2; tests/test_oracle_augment.py, bit-exact).
"""
