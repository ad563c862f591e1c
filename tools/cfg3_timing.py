import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "quick_timing.py")).read().split('if __name__')[0])
run(1024, 1024, 256, [1., 2., 3., 4., 5., 6.], flags=1, tile=(256, 256, 64))
run(2048, 2048, 512, [2., 4., 6.], flags=1, tile=(512, 512, 128))
