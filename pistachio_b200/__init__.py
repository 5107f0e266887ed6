"""pistachio_b200 -- B200-native transfer-matrix hot path with pistachio's Layer/Structure API.

    from pistachio_b200 import transfer_matrix as tm      # drop-in for `from pistachio import transfer_matrix as tm`

The CUDA library (pistachio_b200/csrc -> libpistachio_b200.so) is loaded lazily on
first use; there is no CPU fallback: without the library or without a GPU every
compute call raises.
"""
__version__ = "0.1.0"
