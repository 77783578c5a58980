set -e
cd /root/repo
for f in "-DS3D_KNN_BATCH=1 -DS3D_NN_BATCH=1" "-DS3D_KNN_BATCH=2 -DS3D_NN_BATCH=1" "-DS3D_KNN_BATCH=4 -DS3D_NN_BATCH=1" "-DS3D_KNN_BATCH=1 -DS3D_NN_BATCH=2" "-DS3D_KNN_BATCH=1 -DS3D_NN_BATCH=4" "-DS3D_KNN_BATCH=2 -DS3D_NN_BATCH=2"; do
  bash scripts/variant.sh "$f" --no-chain
done
