"""Map export for visualisation: the step after the hot path in every gradslam example
(`pointclouds.plotly(0).show()`, `pointclouds.open3d(0)`; gradslam/structures/pointclouds.py:1239-1383).

Host-side only: one cloud is copied to numpy (optionally sub-sampled), colours are brought to the range the viewer
expects, and the viewer object is built.  open3d / plotly are imported lazily - they are not needed by the engine and
are absent from the build image - and a missing package raises ImportError at the call, not at import of this package."""
from typing import Optional

import torch

__all__ = ["cloud_arrays", "to_open3d", "to_plotly"]

_HIDDEN_AXIS = dict(showticklabels=False, showgrid=False, zeroline=False, visible=False)


def cloud_arrays(pointclouds, index: int, include_colors: bool = True, max_num_points: Optional[int] = None,
                 include_normals: bool = False, color_range: float = 1.0):
    """(points (n,3) float32, colors (n,3) | None, normals (n,3) | None) as numpy arrays for cloud `index`.
    More than `max_num_points` points are sub-sampled with a random permutation (as the reference does).  Colours are
    returned in [0, color_range]: values above 1.1 are taken to be 0..255 data, anything else 0..1 data."""
    if not isinstance(index, int):
        raise TypeError("Index should be int, but was {}.".format(type(index)))
    points = pointclouds.points_list[index]
    n = points.shape[0]
    keep = None
    if max_num_points is not None and max_num_points < n:
        keep = torch.randperm(n)[:max_num_points].to(points.device)

    def pick(t):
        return (t if keep is None else t[keep]).detach().cpu()

    colors = normals = None
    if include_colors and pointclouds.has_colors:
        c = pick(pointclouds.colors_list[index])
        is_255 = bool((c.max() > 1.1).item()) if c.numel() else False
        if color_range == 1.0:
            c = c / 255 if is_255 else c
        else:
            c = c if is_255 else c * color_range
        colors = torch.clamp(c, min=0.0, max=color_range).numpy()
    if include_normals and pointclouds.has_normals:
        normals = pick(pointclouds.normals_list[index]).numpy()
    return pick(points).numpy(), colors, normals


def to_open3d(pointclouds, index: int, include_colors: bool = True, max_num_points: Optional[int] = None,
              include_normals: bool = False):
    """`open3d.geometry.PointCloud` of cloud `index` (pointclouds.py:1239-1297)."""
    if not isinstance(index, int):
        raise TypeError("Index should be int, but was {}.".format(type(index)))
    try:
        import open3d as o3d
    except ImportError as e:
        raise ImportError("Pointclouds.open3d needs the `open3d` package, which is not installed") from e
    pts, colors, normals = cloud_arrays(pointclouds, index, include_colors, max_num_points, include_normals, 1.0)
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(pts)
    if colors is not None:
        pcd.colors = o3d.utility.Vector3dVector(colors)
    if normals is not None:
        pcd.normals = o3d.utility.Vector3dVector(normals)
    return pcd


def to_plotly(pointclouds, index: int, include_colors: bool = True, max_num_points: Optional[int] = 200000,
              as_figure: bool = True, point_size: int = 2):
    """`plotly.graph_objects.Figure` (or the bare `Scatter3d` with as_figure=False) of cloud `index`
    (pointclouds.py:1299-1383): markers only, uint8 colours, axes hidden."""
    if not isinstance(index, int):
        raise TypeError("Index should be int, but was {}.".format(type(index)))
    try:
        import plotly.graph_objects as go
    except ImportError as e:
        raise ImportError("Pointclouds.plotly needs the `plotly` package, which is not installed") from e
    pts, colors, _ = cloud_arrays(pointclouds, index, include_colors, max_num_points, False, 255.0)
    marker = {"size": point_size}
    if colors is not None:
        marker["color"] = colors.astype("uint8")
    scatter = go.Scatter3d(x=pts[..., 0], y=pts[..., 1], z=pts[..., 2], mode="markers", marker=marker)
    if not as_figure:
        return scatter
    fig = go.Figure(data=[scatter])
    fig.update_layout(showlegend=False, scene=dict(xaxis=dict(_HIDDEN_AXIS), yaxis=dict(_HIDDEN_AXIS),
                                                   zaxis=dict(_HIDDEN_AXIS)))
    return fig
