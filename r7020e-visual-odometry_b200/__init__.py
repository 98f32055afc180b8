"""r7020e-visual-odometry_b200: B200 (sm_100a) hot path of the MATLAB stereo-VO reference.

The directory name is not a valid Python identifier; import it as ``vo_b200`` (the alias package
at the repository root points its ``__path__`` here).

Modules: ``api`` (MATLAB-call mirror over the C ABI), ``vo`` (the VO.m loop), ``synth``
(synthetic stereo frames), ``kitti_eval`` (t_err / r_err), ``shard`` (multi-GPU sharding),
``csrc/`` (CUDA kernels + C ABI, built into ``libvo_b200.so`` next to this file).
"""
from .api import (Context, SIFTPoints, VoError, detectSIFTFeatures, estworldpose,  # noqa: F401
                  extractFeatures, matchFeatures, match_top2, rigidtform3d, sift_batch,
                  triangulate, default_context)
