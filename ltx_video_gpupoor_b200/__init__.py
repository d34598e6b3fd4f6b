"""Import alias: the package directory is `ltx-video-gpupoor_b200/` (not a valid Python identifier),
so this stub points its __path__ there.  `import ltx_video_gpupoor_b200.ops` etc. resolve to files in
that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ltx-video-gpupoor_b200")
__path__.insert(0, _real)
PACKAGE_DIR = _real
