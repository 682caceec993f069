python tools/sweep_fair.py
