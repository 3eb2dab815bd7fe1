import sys; sys.path.insert(0,'/root/repo')
exec(open('/root/repo/tools/dev/regimes.py').read().split("quick = len(sys.argv)")[0])
knobs=[(0,4),(0,6),(0,8),(0,12),(0,16),(0,24)]
case("256x256 + map", 256, 256, 0, 256, 1, True, knobs)
case("640x360 + map", 640, 360, 0, 360, 1, True, knobs)
case("64x64 no map", 64, 64, 0, 64, 1, False, knobs)
case("1280x720 no map", 1280, 720, 0, 720, 1, False, knobs)
