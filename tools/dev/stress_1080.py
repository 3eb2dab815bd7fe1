import sys; sys.path.insert(0, '/root/repo')
exec(open('/root/repo/tools/dev/stress.py').read().split("tot = 0")[0])
stress(1920, 1080, 1, False, 400); stress(1920, 1080, 1, True, 300); stress(336, 141, 7, True, 300)
