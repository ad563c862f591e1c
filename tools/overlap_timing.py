import sys, os
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
sys.argv = ["x", "none"]
exec(open("tools/quick_timing.py").read().split('if __name__')[0])
run(2048, 2048, 512, [2., 4., 6.], flags=1, tile=(512, 512, 128))
run(2048, 2048, 512, [2., 4., 6.], flags=1 | 16, tile=(512, 512, 128))
