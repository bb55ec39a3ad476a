"""blackbox_b200 -- B200-native implementation of BlackBOX's per-frame CCD reduction hot path."""
__version__ = '0.1.0'
