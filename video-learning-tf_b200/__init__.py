"""vlb200: B200-native LRCN hot path behind the video-learning-tf workflow."""
__version__ = "0.2.0"

from . import tfshim  # noqa: E402,F401  (light: numpy only) -- `import vlb200; vlb200.tfshim.install()`
