"""Same entry point as the reference's demo_arterybranch.py:

    python demo_arterybranch.py --cfg config/demo_arterybranch.cfg
"""
import arteryfe as af


def main():
    param = af.ParamParser()          # reads --cfg
    an = af.ArteryNetwork(param)
    an.solve()


if __name__ == '__main__':
    main()
