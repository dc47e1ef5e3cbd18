"""tissue_analysis_b200: the per-label voxel-scan hot path of VirtualPlants/tissue_analysis
(SpatialImageAnalysis3D feature extractors) on hand-written sm_100a CUDA kernels behind a C ABI."""
from .serial import imread, imsave
from .spatial_image import SpatialImage
from .spatial_image_analysis import (NPLIST, LIST, DICT, AbstractSpatialImageAnalysis, SpatialImageAnalysis,
                                     SpatialImageAnalysis3D)

__all__ = ["SpatialImage", "imread", "imsave", "SpatialImageAnalysis", "SpatialImageAnalysis3D", "AbstractSpatialImageAnalysis",
           "NPLIST", "LIST", "DICT"]
__version__ = "0.1.0"
