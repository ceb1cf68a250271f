"""B200-native batched simulator for the per-env step() of RoboticsPlayroomPybullet's
gym envs (UR5Reach-v0, pandaPick-v0, UR5PlayAbsRPY1Obj-v0).  See DESIGN.md."""
__version__ = '0.1.0'
