"""traffic_b200 - B200-native per-timestep vehicle update for donjpierce/traffic.

  native    ctypes binding of libtraffic_b200.so (include/traffic_b200.h) + DeviceSim
  facade    drop-in Cars / TrafficLights (the reference's cars.py API); dropin/cars.py makes it importable as `cars`
  init      drop-in initialisers over GPU-batched routing (routing)
  ensemble  replica ensembles (Env rollouts) sharded over GPUs
  state     DataFrame <-> structure-of-arrays packing
  synth     synthetic road worlds (no network, no osmnx)
"""
__version__ = "0.1.0"
