"""ld_tools_b200 -- B200-native engine for ld-tools' LD hot path (calc_ld as driven by ld_lite,
ld_area and ld_triangle).  See DESIGN.md; the C ABI is include/ldx.h."""
from .calc_ld import calc_ld
from .engine import Context, HostText, Store, LdxError

__all__ = ["calc_ld", "Context", "HostText", "Store", "LdxError"]
__version__ = "0.1.0"
