import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import bench_configs as b
for G in (4,8,16,32):
    try: b.run(f"C3 G={G}", "c3_arm.hsrb", 16384, pan=True, steps=2, warmup=1, lanes=G)
    except Exception as e: print("C3",G,"ERR",e)
for G in (16,32):
    try: b.run(f"C5 G={G}", "c5_clutter.hsrb", 4096, pan=False, steps=2, warmup=1, lanes=G)
    except Exception as e: print("C5",G,"ERR",e)
