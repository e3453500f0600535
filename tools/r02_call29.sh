#!/bin/bash
timeout 600 python tools/sweep_tile_n.py 2>&1 | tail -30
