"""No-op stand-in for pyecharts (visualisation is out of scope)."""
