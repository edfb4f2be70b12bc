python -c "import torch; torch.cuda.init()" 
bash tools/driver_startup_probe.sh
