"""improver_b200 — B200-native (sm_100a) replacement for the gridded
neighbourhood-processing hot path of metoppv/improver.

The plugin surface mirrors the reference (`improver.nbhood.nbhood`,
`improver.threshold`, `improver.utilities.neighbourhood_tools`); the arithmetic
runs in hand-written CUDA kernels reached through a C ABI
(include/improver_b200.h).  There is no CPU fallback.
"""
from typing import Any

__version__ = "0.1.0"


class BasePlugin:
    """Mirror of improver.BasePlugin (improver/__init__.py:24-50): calling the
    plugin runs `process`."""

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        return self.process(*args, **kwargs)

    def process(self, *args: Any, **kwargs: Any) -> Any:   # pragma: no cover - abstract
        raise NotImplementedError


class PostProcessingPlugin(BasePlugin):
    """Mirror of improver.PostProcessingPlugin (improver/__init__.py:53-89):
    results with a non-default title get a "Post-Processed" prefix."""

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        result = super().__call__(*args, **kwargs)
        self.post_processed_title(result)
        return result

    @staticmethod
    def post_processed_title(cube) -> None:
        attrs = getattr(cube, "attributes", None)
        if not isinstance(attrs, dict):
            return
        default_title = "unknown"
        title = attrs.get("title")
        if title and title != default_title and "Post-Processed" not in title:
            attrs["title"] = f"Post-Processed {title}"
