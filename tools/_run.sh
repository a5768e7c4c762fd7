bash tools/final_profile.sh r2
