"""vlb200: B200-native LRCN hot path behind the video-learning-tf workflow."""
__version__ = "0.1.0"
