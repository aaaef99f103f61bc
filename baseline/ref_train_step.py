"""The reference's training iteration, statement by statement (baseline arm / drop-in proof; NOT product).

`torch/train.py:399-757` past `num_iters_geo_only`, with its default switches except: GAN and VGG style terms off
(`--weight_disc_loss 0`, the style weights' defaults) and the 2D semantic branch (`--pred_3d_semantic ''`, train.py:740-747),
i.e. BASELINE.json configs[3].  Everything executed is the reference's own code: `model.Generator`, the `loss` module and
`utils/raycast_rgbd/raycast_rgbd.py` are imported unmodified from `baseline/_ref` (baseline/ref_loader.py); only the
statements of train.py's loop body are restated here (cited line by line), because train.py itself is a script around a
dataset that does not ship.  The native module behind the reference's wrapper is selectable:

    native="reference"  the compiled reference extension (oracle/_ref)                 -> the baseline arm
    native="ours"       this repository's drop-in `raycast_rgbd_cuda` over the C ABI   -> the drop-in proof: the reference's
                        wrapper, losses and loop run unchanged on the new kernels
"""
import torch
import torch.nn.functional as F

from . import ref_loader


class RefTrainStep:
    def __init__(self, model, batch_size, dims3d, width, height, class_weight, native="reference", voxelsize=0.02,
                 truncation=3.0, weight_occ_loss=1.0, weight_sdf_loss=0.1, weight_depth_loss=1.0, weight_color_loss=1.0,
                 weight_semantic_loss=0.1, weight_surf_geo=1.0, weight_missing_geo=5.0, logweight_sdf=True,
                 max_num_locs_per_sample=640000):
        self.model = model
        self.loss_util = ref_loader.load_module("loss")
        wrapper = ref_loader.load_wrapper(native)
        self.batch_size, self.dims3d = batch_size, tuple(dims3d)
        self.voxelsize, self.truncation = voxelsize, truncation
        self.w = dict(occ=weight_occ_loss, sdf=weight_sdf_loss, depth=weight_depth_loss, color=weight_color_loss,
                      semantic=weight_semantic_loss)
        self.weight_surf_geo, self.weight_missing_geo, self.logweight_sdf = weight_surf_geo, weight_missing_geo, logweight_sdf
        self.class_weight = class_weight
        ray_increment = 0.3 * truncation            # train.py:134
        thresh_sample_dist = 50.5 * ray_increment   # train.py:135
        self.raycaster_rgbd = wrapper.RaycastRGBD(batch_size, self.dims3d, width, height, depth_min=0.1 / voxelsize,
                                                  depth_max=6.0 / voxelsize, thresh_sample_dist=thresh_sample_dist,
                                                  ray_increment=ray_increment,
                                                  max_num_locs_per_sample=max_num_locs_per_sample)  # train.py:138-141
        self.last = {}

    def __call__(self, sample, optimizer=None):
        loss_util, args_truncation = self.loss_util, self.truncation
        raycaster_rgbd = self.raycaster_rgbd
        inputs, mask, known = sample["input"], sample["mask"], sample.get("known")
        target_for_sdf, target_for_colors = loss_util.compute_targets(sample["sdf"], args_truncation, True, known,
                                                                      sample["colors"])                       # :448
        target_for_semantics = sample["semantics"]
        if optimizer is not None:
            optimizer.zero_grad()                                                                             # :461
        output_occ, output_sdf, output_color, output_semantic = self.model(
            inputs, mask, pred_sdf=[True, True], pred_color=True, pred_semantic=True)                        # :465
        loss = 0.0
        input_occ = torch.abs(inputs[:, :1]) < (args_truncation - 0.01)                                       # :475
        weight = loss_util.compute_dense_geo_weights(target_for_sdf, input_occ, args_truncation, self.weight_surf_geo,
                                                     self.weight_missing_geo)                                 # :477
        empty = torch.nn.Sigmoid()(output_occ.detach()) < 0.5                                                 # :480
        weight[empty] = 0
        loss_occ = loss_util.compute_geo_occ_loss(target_for_sdf, output_occ, known, weight, args_truncation)  # :482
        loss += self.w["occ"] * loss_occ
        loss_sdf = loss_util.compute_geo_loss(target_for_sdf, None, output_sdf, known, weight, self.logweight_sdf)  # :489
        loss += self.w["sdf"] * loss_sdf
        locs = torch.nonzero((torch.abs(output_sdf.detach()[:, 0]) < args_truncation) & ~empty[:, 0])         # :495
        locs = torch.cat([locs[:, 1:], locs[:, :1]], 1)                                                       # :498
        output_sdf = [locs, output_sdf[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]]                   # :499
        output_color = [locs, output_color[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]]               # :505
        output_semantic = output_semantic[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]                 # :508
        self.last = dict(num_locs=len(output_sdf[0]), loss_occ=loss_occ.detach(), loss_sdf=loss_sdf.detach())
        if 0 < len(output_sdf[0]) <= raycaster_rgbd.get_max_num_locs_per_sample() * self.batch_size:         # :524-529
            images_color = sample["images_color"]
            images_depth = sample["images_depth"]
            intrinsics = sample["images_intrinsic"]
            view_matrix = sample["view_matrix"]                                                               # :534
            images_depth = images_depth.unsqueeze(1)                                                          # :536
            output_normals = loss_util.compute_normals_sparse(output_sdf[0], output_sdf[1], target_for_sdf.shape[2:],
                                                              transform=torch.inverse(view_matrix))           # :542
            # input raycast (:556-578)
            input_locs = torch.nonzero(torch.abs(inputs[:, 0]) < args_truncation)
            input_locs = torch.cat([input_locs[:, 1:], input_locs[:, :1]], 1)
            input_vals = inputs[input_locs[:, -1], :, input_locs[:, 0], input_locs[:, 1], input_locs[:, 2]]
            input_normals = loss_util.compute_normals(inputs[:, :1], input_locs, transform=torch.inverse(view_matrix))
            raycast_color, _, raycast_normal, _ = raycaster_rgbd(
                input_locs, input_vals[:, :1].contiguous(), input_vals[:, 1:4].contiguous(), input_normals, None,
                view_matrix, intrinsics)
            invalid = raycast_color == -float('inf')
            input2d = raycast_color.clone() * 2 - 1
            input2d[invalid] = 0
            normals = raycast_normal.clone()
            invalid = raycast_normal == -float('inf')
            normals[invalid] = 0
            input2d = torch.cat([input2d, normals], 3).permute(0, 3, 1, 2).contiguous()
            # target raycast (:581-622)
            locs = torch.nonzero(torch.abs(target_for_sdf[:, 0]) < args_truncation)
            locs = torch.cat([locs[:, 1:], locs[:, :1]], 1).contiguous()
            vals = target_for_sdf[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]].contiguous()
            colors = target_for_colors[locs[:, -1], locs[:, 0], locs[:, 1], locs[:, 2], :].float() / 255.0
            target_normals = loss_util.compute_normals_sparse(locs, vals, target_for_sdf.shape[2:],
                                                              transform=torch.inverse(view_matrix))
            target_semantics = target_for_semantics[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]
            target_semantics_onehot = F.one_hot(target_semantics[:, 0].long(), 15)[..., :-1].float().contiguous()
            raycast_color, _, raycast_normal, raycast_semantic = raycaster_rgbd(
                locs, vals, colors.contiguous(), target_normals, target_semantics_onehot, view_matrix, intrinsics)
            cat = torch.cat((raycast_semantic.clone(), torch.ones(raycast_semantic.shape[:-1] + (1,),
                                                                  device=raycast_semantic.device)), dim=-1)
            _, target2d_label = torch.max(cat, dim=-1, keepdim=True)                                          # :615
            target2d_label = target2d_label.to(torch.uint8)
            color = (output_color[1] + 1) * 0.5                                                               # :618
            semantic = output_semantic.clone()                                                                # :622
            # prediction raycast (:626-627)
            raycast_color, raycast_depth, raycast_normal, raycast_semantic = raycaster_rgbd(
                output_sdf[0], output_sdf[1], color, output_normals, semantic, view_matrix, intrinsics)
            raycast_depth = raycast_depth.unsqueeze(1) * self.voxelsize                                       # :635
            valid = (raycast_depth != -float('inf')) & (images_depth != 0)
            loss_depth = torch.mean(torch.abs(raycast_depth[valid] - images_depth[valid]))
            loss += self.w["depth"] * loss_depth
            loss_color = loss_util.compute_2dcolor_loss(raycast_color, images_color.permute(0, 2, 3, 1), None)  # :643
            loss += self.w["color"] * loss_color
            valid = torch.logical_and(target2d_label[..., 0] < 14, raycast_semantic[..., 0] != -float('inf'))   # :744
            loss_semantic = F.cross_entropy(raycast_semantic[valid].view(-1, raycast_semantic.shape[-1]),
                                            target2d_label[valid].view(-1).long(), weight=self.class_weight)
            loss += loss_semantic * self.w["semantic"]
            self.last.update(terms2d=torch.stack([loss_depth.detach(), loss_color.detach(), loss_semantic.detach()]),
                             target2d_label=target2d_label[..., 0])
        if optimizer is not None:
            loss.backward()                                                                                   # :756
            optimizer.step()                                                                                  # :757
        return loss.detach() if torch.is_tensor(loss) else torch.tensor(loss)
