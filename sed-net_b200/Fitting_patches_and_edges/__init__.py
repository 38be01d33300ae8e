"""Drop-in replacements for the stage-2 hot-path modules of the reference's ``Fitting_patches_and_edges`` directory
(SURVEY.md section 8f row 2): ``primitive_forward_v2.Fit``, ``proj_2_edge_utils`` (instance adjacency maps) and
``pointnet2.pointnet2_utils.three_nn``, every call running on the sm_100a kernels of libsednet_b200.so (no CPU
fallback), plus ``wire`` -- the text files stage 1 hands to stage 2 (generate_predictions_aug.py:424-437)."""
