D=/tmp/udh_sup; rm -rf $D; mkdir -p $D
python homography_CNN_synthetic.py --mode train --loss_type h_loss --batch_size 128 --synthetic 65536 --num_gpus 1 --numeric bf16x3 --lr 1e-4 \
    --num_total_steps 12000 --model_dir $D/m --log_dir $D/l --results_dir $D/r 2>&1 | awk '/step [0-9]*000,/' | cut -c1-200
python homography_CNN_synthetic.py --mode test --loss_type h_loss --batch_size 128 --synthetic 2048 --num_gpus 1 --numeric bf16x3 --do_augment 0 \
    --model_dir $D/m --log_dir $D/l --results_dir $D/r 2>&1 | tail -4 | cut -c1-200
