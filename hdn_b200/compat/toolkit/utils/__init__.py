from .statistics import (alignment_error, overlap_ratio, polygon_centroid, polygon_is_simple, success_4pts_error, success_centroid_error,  # noqa: F401
                         success_error, success_overlap)
