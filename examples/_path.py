"""Puts the drop-in package (``gym-acas2d_b200/``) on sys.path for the example scripts."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-acas2d_b200"))
