"""TEST INFRASTRUCTURE - the CPU oracle for the vehicle-update path. Never imported by traffic_b200."""
