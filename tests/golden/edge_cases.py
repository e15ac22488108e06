"""Seeded inputs of the PointFusion edge cases (shared by make_golden_params.py, which freezes the reference's outputs
for them, and by the CPU / GPU tests, which rebuild the same inputs).  Depends on gradslam_b200.synthetic only."""
from gradslam_b200.synthetic import make_sequence

EDGE_CASES = ["edge_empty_mid_frame", "edge_empty_first_frame", "edge_empty_element", "edge_half_frames",
              "edge_no_overlap"]


def edge_inputs(name):
    rgb, depth, K, poses = make_sequence(2, 4, 48, 64, seed=21)
    if name == "edge_empty_mid_frame":      # an all-invalid frame in the middle of every sequence
        depth[:, 2] = 0
    elif name == "edge_empty_first_frame":  # one sequence starts with an all-invalid frame (its map starts later)
        depth[1, 0] = 0
    elif name == "edge_empty_element":      # one sequence never sees a valid depth: its map stays empty
        depth[1] = 0
    elif name == "edge_half_frames":        # later frames only cover the left half of the image
        depth[:, 1:, :, 32:] = 0
    elif name == "edge_no_overlap":         # frame 2 looks at nothing the map contains: no correspondences at all
        poses[:, 2, :3, 3] += 5.0
    else:
        raise KeyError(name)
    return rgb, depth, K, poses
