#!/usr/bin/env python
"""Generates tests/golden/remap_c8_20x10_order{1,2}.nc: what the UNMODIFIED reference's setup_conserve_interp WRITE branch
(conserve_interp.c:368-443) hands its netCDF layer for C8 -> 20x10 (recorded by oracle/shim/io_stubs.c), rendered per the classic
netCDF format specification (64-bit offset) by tests/test_remap_cpu.py::_classic_bytes.  Needs /root/reference (oracle/_ref).
   python tests/golden/make_remap_golden.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import xgtest  # noqa: E402
from test_remap_cpu import _classic_bytes  # noqa: E402

pkg = xgtest.package()
lonc, latc = pkg.cubed_sphere_grid(8)
lon2, lat2 = pkg.latlon_grid(20, 10)
for order, op in ((1, xgtest.ORDER1), (2, xgtest.ORDER2)):
    xgtest.ref_setup(lonc, latc, lon2, lat2, op, remap=("golden.nc", 1))
    dims, variables = xgtest.ref_store_file("golden.nc")
    path = os.path.join(HERE, f"remap_c8_20x10_order{order}.nc")
    open(path, "wb").write(_classic_bytes(2, dims, variables))
    print(path, os.path.getsize(path), "bytes,", dims[1][1], "cells")
