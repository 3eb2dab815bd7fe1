import sys; sys.path.insert(0,'/root/repo')
exec(open('/root/repo/tools/dev/seg_sweep3.py').read().split("for seg in [0,443")[0])
for (W,H,F,mp,it) in [(3840,2160,64,True,10),(1920,1080,512,True,6),(16384,16384,1,True,10),(16384,2058,1,True,20),(16384,4106,1,True,20),(16384,8202,1,True,20),(3840,2160,1,True,30),(1920,1080,1,False,30),(1920,1080,64,True,10),(256,256,1,True,30),(7680,4320,1,True,20),(1280,720,16,True,20)]:
    print(W,H,F,mp,"auto",run(W,H,F,mp,0,it), flush=True)
