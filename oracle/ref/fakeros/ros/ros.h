/* oracle/ref/fakeros -- CPU ORACLE, TEST INFRASTRUCTURE ONLY: the reference sources include ROS / profiler headers they do not need on the cuboid path; empty stand-ins. */
