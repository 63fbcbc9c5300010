#!/usr/bin/env python3
"""The latency-sized entry points of bench.py's `candidate_loops` alone (SearchByBoW, frame build, SearchByProjection), for
`ncu --metrics gpu__time_duration.sum` launch lists: python tools/profile_loops.py"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

orb = importlib.import_module("cooperative-orb-slam_b200")
synth = importlib.import_module("cooperative-orb-slam_b200.synth")
print(json.dumps(bench.bench_candidate_loops(orb, synth, 0)))
