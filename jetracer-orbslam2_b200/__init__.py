"""B200-native ORB front-end."""
