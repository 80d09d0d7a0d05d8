"""Same entry point as the reference's demo_arterybranch.py:

    python demo_arterybranch.py --cfg config/demo_arterybranch.cfg

With the configs as shipped and HEAD's hard-coded Windkessel factors (0.3, 0.01;
artery_network.py:408-409) the artery Newton fails in step 0, as it does in the
reference.  ``--wk-scale 1,1`` (not in the reference) runs the same
config with the R1, R2 the file states.
"""
import argparse

import arteryfe as af


def main():
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument('--wk-scale', default=None, help="R1,R2 factors instead of HEAD's 0.3,0.01")
    extra = ap.parse_known_args()[0]
    wk_scale = tuple(float(v) for v in extra.wk_scale.split(',')) if extra.wk_scale else None
    param = af.ParamParser()          # reads --cfg
    an = af.ArteryNetwork(param, wk_scale=wk_scale)
    an.solve()


if __name__ == '__main__':
    main()
