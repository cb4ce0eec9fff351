"""Host-side Python mirror of the C-ABI in include/smj.h (the product is libsmj.so; the reference's own
host is C -- see host/app.c for the drop-in driver).  Import as ``smj_b200`` through the repo-root shim,
because the directory name carries hyphens."""
from .smj import (  # noqa: F401
    SmjError, Config, Stats, Table, lib, build, lib_path, select, sort, merge, join, join_count, run,
    device_table, synth_device_table, free, JOIN_ZIP, JOIN_MANY, csv_parse, csv_format, from_i64, to_i64, to_numpy,
)
from . import datagen  # noqa: F401
from . import dist  # noqa: F401
