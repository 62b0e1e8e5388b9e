"""samcon_b200 -- B200-native data-parallel hot path of SamCon behind the reference's
search-loop / env interface.  See DESIGN.md."""
__version__ = "0.1.0"
