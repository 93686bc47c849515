"""TEST INFRASTRUCTURE - probe of the reference's RL featuriser (navigation.StateView, navigation.py:131-418) at HEAD,
run in the build container on the unmodified reference: what `Env.reset()` (environment.py:43-54) would get for the first
agents of a seeded 59-car world.  Recorded in profiles/r2_ref_stateview_probe.txt; it is why SURVEY.md 8(f) row 4 has no
parity target."""
import random
import sys
import traceback
import warnings

warnings.filterwarnings("ignore")
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ref_harness as rh  # noqa: E402

ref_cars, ref_sim = rh.install_reference()
import navigation as nav  # noqa: E402
from traffic_b200 import synth  # noqa: E402

g = synth.SynthGraph(nx_=12, ny_=12, x0=566000.0, y0=4186000.0, seed=3)
random.seed(3)
cars = ref_sim.init_random_node_start_location(60, g)
lights = ref_sim.init_traffic_lights(g, prescale=2)
C = ref_cars.Cars(cars, g)
L = ref_cars.TrafficLights(lights, g)
for agent in range(12):
    try:
        sv = nav.StateView(graph=g, car_index=agent, cars=C.state, lights=L.state)
        print("agent", agent, "route nodes", len(sv.route), "state", sv.determine_state()[0])
    except Exception as e:  # noqa: BLE001
        fr = traceback.extract_tb(e.__traceback__)[-1]
        where = [f for f in traceback.extract_tb(e.__traceback__) if "navigation.py" in f.filename][-1]
        print("agent", agent, "route nodes", len(cars.loc[agent]["route"]), "->", type(e).__name__, str(e)[:60],
              "| navigation.py:%d in %s" % (where.lineno, where.name))
