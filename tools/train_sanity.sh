# End-to-end sanity of the CLI in the certified two-limb mode: supervised (h_loss) and unsupervised (l1_loss) training on
# on-device synthetic pairs with the fused augmentation, TF-checkpoint save, then --mode test restoring that checkpoint.
set -e
D=/tmp/udh_sanity; rm -rf $D; mkdir -p $D
for lt in h_loss l1_loss; do
  echo "##### --loss_type $lt"
  python homography_CNN_synthetic.py --mode train --loss_type $lt --batch_size 128 --synthetic 65536 --num_gpus 1 --numeric bf16x3 --lr $([ $lt = h_loss ] && echo 1e-4 || echo 5e-4) \
    --num_total_steps ${STEPS:-6000} --model_dir $D/m_$lt --log_dir $D/l_$lt --results_dir $D/r_$lt 2>&1 | awk 'NR<=3 || /step [0-9]*000,/ || /checkpoint|saved|Saved/' | cut -c1-260
  ls $D/m_$lt | head -8
  python homography_CNN_synthetic.py --mode test --loss_type $lt --batch_size 128 --synthetic 2048 --num_gpus 1 --numeric bf16x3 --do_augment 0 \
    --model_dir $D/m_$lt --log_dir $D/l_$lt --results_dir $D/r_$lt 2>&1 | tail -4 | cut -c1-260
done
