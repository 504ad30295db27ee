timeout 300 python tools/profile_unet.py 8 bench 2>&1 | cut -c1-60,150-215 | sed -n 3,16p; 
timeout 300 python tools/profile_unet.py 8 bench 2>&1 | grep "Self CUDA time total"
