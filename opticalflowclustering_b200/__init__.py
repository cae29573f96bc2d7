"""B200-native (sm_100a) hot path of menmitsu/opticalFlowClustering's
k-means-color-clustering pipeline: Farneback flow -> HSV visualisation ->
grid-cell aggregation -> k-means -> cosine similarity.

The operators need the in-tree CUDA library ``libofc.so`` (built by
``__graft_entry__.build()``); there is no CPU fallback.
"""
__all__ = ["flow", "grid", "pipeline", "synthetic", "kmeans", "cosine", "computeOpticalFlowModule", "computeOpticalFlow",
           "KmeanGrids", "drawGridsAndOutputCSV", "drawGridsAndOutputCSVChange", "color_kmeans", "color_kmeansChange", "findCosineDifferentVectors", "computeVectorDistance"]
