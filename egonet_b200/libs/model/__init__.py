"""Mirror of the reference's ``libs.model`` package (upstream ``libs/model/__init__.py``
exposes ``heatmapModel`` so that ``eval('models.heatmapModel.hrnet.get_pose_net')``
in ``egonet.py:43-44`` resolves)."""
from . import heatmapModel  # noqa: F401
