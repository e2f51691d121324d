"""B200-native dense geometric core of textureless-3d-reconstruction.

Drop-in surfaces (same names/signatures as the reference):
  depth_to_reconstruction      ReconstructionConfig, DepthImageLoader, DenseReconstructor,
                               DepthToReconstructionPipeline, main
  depth_enhanced_reconstruction CameraIntrinsics, DensePointCloudGenerator, DepthEnhancedReconstruction
  depth_processor              CameraIntrinsics (+from_json), PointCloudGenerator
Device API: runtime.Context, runtime.TSDFVolume.  C ABI: include/t3d.h (libt3d.so).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "runtime", "depth_to_reconstruction", "depth_enhanced_reconstruction",
           "depth_processor", "synthetic"]
__version__ = "0.1.0"
