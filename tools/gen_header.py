"""Regenerate include/prb_model.h from the Python schema (single source of truth)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from roboticsplayroompybullet_b200.model import c_header
if __name__ == '__main__':
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'include', 'prb_model.h')
    open(p, 'w').write(c_header())
    print('wrote', p)
